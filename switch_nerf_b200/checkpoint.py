"""Checkpoint interop for the hot-path model (reference models/model_utils.py:12-67, 136-151).

The reference stores expert parameters in one of two layouts:
  expertmlp  : layers.L.experts.0.weights.J [E, in, out], layers.L.experts.0.bias.J [E, 1, out]   (training)
  seqexperts : layers.L.experts.0.experts.E.layers.J.weight [out, in], ....bias [out]             (--expertmlp2seqexperts)
and DDP checkpoints carry a `module.` prefix (stripped at model_utils.py:147).  The CUDA path packs the expertmlp
layout; these helpers accept either and emit either.  Host-side tensor bookkeeping only."""
import re
from collections import OrderedDict
from typing import Dict

import torch
from torch import Tensor

_SEQ = re.compile(r"^(?P<pre>.*layers\.(?P<L>\w+)\.experts\.0\.)experts\.(?P<e>\d+)\.layers\.(?P<j>\d+)\.(?P<kind>weight|bias)$")
_MLP = re.compile(r"^(?P<pre>.*layers\.(?P<L>\w+)\.experts\.0\.)(?P<kind>weights|bias)\.(?P<j>\d+)$")


def strip_module_prefix(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """consume_prefix_in_state_dict_if_present(state_dict, 'module.') (model_utils.py:147), non-destructive."""
    return OrderedDict((k[len("module."):] if k.startswith("module.") else k, v) for k, v in sd.items())


def to_expertmlp(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """seqexperts keys -> stacked expertmlp tensors (inverse of convert_to_seqexperts, model_utils.py:12-28).
    Keys that are not per-expert Linear parameters pass through; `module.` prefixes are removed."""
    sd = strip_module_prefix(sd)
    out, groups = OrderedDict(), {}
    for k, v in sd.items():
        m = _SEQ.match(k)
        if m is None:
            out[k] = v
            continue
        groups.setdefault((m["pre"], int(m["j"]), m["kind"]), {})[int(m["e"])] = v
    for (pre, j, kind), per_expert in groups.items():
        n = len(per_expert)
        if sorted(per_expert) != list(range(n)):
            raise KeyError(f"{pre}: expert ids of layer {j} are not 0..{n - 1}")
        if kind == "weight":      # [out, in] per expert -> [E, in, out]
            out[f"{pre}weights.{j}"] = torch.stack([per_expert[e].t() for e in range(n)], 0).contiguous()
        else:                     # [out] -> [E, 1, out]
            out[f"{pre}bias.{j}"] = torch.stack([per_expert[e].reshape(1, -1) for e in range(n)], 0).contiguous()
    return out


def to_seqexperts(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """expertmlp -> seqexperts keys (what --expertmlp2seqexperts produces, minus the `module.` prefix)."""
    out = OrderedDict()
    for k, v in strip_module_prefix(sd).items():
        m = _MLP.match(k)
        if m is None:
            out[k] = v
            continue
        for e, t in enumerate(torch.unbind(v, 0)):
            if m["kind"] == "weights":
                out[f"{m['pre']}experts.{e}.layers.{m['j']}.weight"] = t.t().contiguous()
            else:
                out[f"{m['pre']}experts.{e}.layers.{m['j']}.bias"] = t.reshape(-1).contiguous()
    return out


def load_checkpoint(model: torch.nn.Module, ckpt, weight_key: str = "model_state_dict", map_location="cpu"):
    """model_utils.py:136-151: load `ckpt[weight_key]` (path or dict; a bare state_dict is accepted too), strip
    `module.`, accept either expert layout, keep parameters the checkpoint does not name."""
    if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    sd = ckpt[weight_key] if isinstance(ckpt, dict) and weight_key in ckpt else ckpt
    merged = model.state_dict()
    merged.update(to_expertmlp(sd))
    return model.load_state_dict(merged)
