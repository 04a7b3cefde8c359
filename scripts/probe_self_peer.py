"""GPU probe: does the self-peer engine (world = 1) get symmetric memory + an NVLS multicast mapping from torch's allocator?
Prints one JSON line; used once to decide how the world = 1 multimem tests are gated."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BAGUA_SELF_PEER"] = "1"
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29871")
import torch  # noqa: E402

import bagua_b200 as bagua  # noqa: E402

torch.cuda.set_device(0)
bagua.init_process_group()
eng = bagua.communication._get_default_group().peer_engine()
out = {"engine": eng is not None}
if eng is not None:
    out.update(world=eng.world, multicast=bool(eng.has_multicast), symm=bool(eng.self_peer_symm))
    sl = eng.alloc(1 << 20)
    out["slice_multicast"] = bool(sl.has_multicast)
    if sl.has_multicast:
        from bagua_b200.core import native

        t = sl.view(torch.float32)
        t.fill_(3.0)
        op, chosen = eng.make_allreduce_op(sl, sl, 1 << 20, torch.float32, True, "multimem")
        native().run_op(op, torch.cuda.current_stream().cuda_stream, 0)
        torch.cuda.synchronize()
        out["multimem_variant"] = chosen
        out["multimem_ok"] = bool((t == 3.0).all().item())
        out["error_code"] = eng.comm.error_code()
print(json.dumps(out))
