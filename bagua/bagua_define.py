from bagua_b200.define import *  # noqa: F401,F403
from bagua_b200.define import BaguaCoreTelemetrySpan, BaguaHyperparameter, TensorDeclaration, TensorDtype, get_tensor_declaration_bytes  # noqa: F401
