"""Model zoo of the BASELINE configs: small MNIST CNN, VGG16, ResNet-50, BERT-large (QA head), GPT-2-medium MoE-8.
Architectures are the standard ones (random init — no checkpoints are shipped); they exist so that benchmarks and
examples run without torchvision / transformers."""
from .mnist import MnistNet  # noqa: F401
from .vgg import vgg16, VGG  # noqa: F401
from .resnet import resnet50, ResNet  # noqa: F401
from .bert import BertConfig, BertForQuestionAnswering, bert_large_config, bert_qa_from_pretrained, convert_hf_state_dict, save_pretrained, to_hf_state_dict  # noqa: F401
from .gpt2_moe import GPT2MoEConfig, GPT2MoE, gpt2_medium_moe8_config  # noqa: F401

_REGISTRY = {"vgg16": vgg16, "resnet50": resnet50, "mnist": MnistNet}


def get_model(name: str, **kwargs):
    """A model by name: the built-in zoo first, then any ``torchvision.models`` architecture (random init) when torchvision
    is installed — what the reference's synthetic benchmark does with ``getattr(torchvision.models, name)()``."""
    if name in _REGISTRY:
        return _REGISTRY[name](**kwargs)
    try:
        import torchvision.models as tvm
    except ImportError as e:
        raise KeyError(f"unknown model {name!r}; built in: {sorted(_REGISTRY)} (torchvision is not installed)") from e
    ctor = getattr(tvm, name, None)
    if not callable(ctor):
        raise KeyError(f"unknown model {name!r}; built in: {sorted(_REGISTRY)}, or any torchvision.models architecture")
    return ctor(weights=None, **kwargs)
