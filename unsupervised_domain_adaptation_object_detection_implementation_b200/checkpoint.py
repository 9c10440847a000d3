"""Checkpoint files in the layout the reference writes and reads (mmcv.runner.save_checkpoint / load_checkpoint as
used by tools/DA_train.py:258-263 and mmdet/apis/train.py:167,202): a torch.save'd dict
{'meta': {...}, 'state_dict': OrderedDict[str, cpu tensor], 'optimizer': {...}?}.  Parameter names of the DA modules
are the reference's (tests/golden/state_dict_surface.json), so files move in both directions.

Sharded optimizers: pass `peer_optimizer=` so that the fp32 master (current only on each rank's own slice,
peer.PeerShardedSGD) is gathered before the state_dict is taken; only rank 0 writes."""
import collections
import os
import re
import time

import torch


def weights_to_cpu(state_dict):
    out = collections.OrderedDict()
    for k, v in state_dict.items():
        out[k] = v.detach().cpu() if torch.is_tensor(v) else v
    return out


def save_checkpoint(model, filename, optimizer=None, meta=None, peer_optimizer=None, rank=0):
    if peer_optimizer is not None:
        peer_optimizer.gather_master()            # collective: every rank calls save_checkpoint
    if rank != 0:
        return None
    meta = dict(meta or {})
    meta.setdefault("time", time.asctime())
    module = model.module if hasattr(model, "module") else model          # unwrap (MM)DataParallel
    if hasattr(module, "CLASSES") and module.CLASSES is not None:
        meta.setdefault("CLASSES", module.CLASSES)
    ckpt = {"meta": meta, "state_dict": weights_to_cpu(module.state_dict())}
    if optimizer is not None and hasattr(optimizer, "state_dict"):
        ckpt["optimizer"] = optimizer.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    tmp = f"{filename}.tmp.{os.getpid()}"
    torch.save(ckpt, tmp)
    os.replace(tmp, filename)                     # a reader never sees a partial file
    return ckpt


def load_checkpoint(model, filename, map_location="cpu", strict=False, revise_keys=((r"^module\.", ""),)):
    """Returns the checkpoint dict.  Accepts a full checkpoint or a bare state_dict; `revise_keys` are (regex, replacement)
    pairs applied to every key (default: strip the DataParallel prefix), as in mmcv.  Non-strict loading reports missing
    and unexpected keys in the returned dict under 'load_report'."""
    ckpt = torch.load(filename, map_location=map_location, weights_only=False)
    if not isinstance(ckpt, dict):
        raise RuntimeError(f"No state_dict found in checkpoint file {filename}")
    state = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    revised = collections.OrderedDict()
    for k, v in state.items():
        for pat, rep in revise_keys:
            k = re.sub(pat, rep, k)
        revised[k] = v
    module = model.module if hasattr(model, "module") else model
    result = module.load_state_dict(revised, strict=strict)
    for p in module.parameters():                 # cached bf16 operand copies are stale after an in-place load
        if hasattr(p, "_da_shadow"):
            del p._da_shadow
    out = ckpt if "state_dict" in ckpt else {"state_dict": state, "meta": {}}
    out["load_report"] = {"missing_keys": list(result.missing_keys), "unexpected_keys": list(result.unexpected_keys)}
    return out
