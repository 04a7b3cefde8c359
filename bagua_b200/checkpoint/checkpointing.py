"""MoE-aware checkpoint save / load (reference: bagua/torch_api/checkpoint/checkpointing.py:1-363).

On-disk layout is the reference's (Megatron/DeepSpeed style) so checkpoints are interchangeable::

    <path>/latest_checkpointed_iteration.txt
    <path>/iter_0000042/mp_rank_00_model_states.pt                         # rank 0: non-expert model, iteration, lr scheduler (+ optimizer when no MoE)
    <path>/iter_0000042/expert_<global id>_mp_rank_00_model_states.pt      # every rank: each of its local experts, renamed to global ids
    <path>/iter_0000042/expert_parallel_rank_<r>_mp_rank_00_optim_states.pt  # every rank: its optimizer state (MoE only)
"""
from __future__ import annotations

import logging
import os
import re
import sys
from collections import defaultdict
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["save_checkpoint", "load_checkpoint"]

logger = logging.getLogger(__name__)
_MOE_PREFIX = ".bagua_moe.experts.bagua_experts."


def _iter_dir(path: str, iteration: int, release: bool = False) -> str:
    return os.path.join(path, "release" if release else f"iter_{iteration:07d}")


def _get_model_ckpt_name(path, iteration, mp_rank=0, release=False) -> str:
    return os.path.join(_iter_dir(path, iteration, release), f"mp_rank_{mp_rank:02d}_model_states.pt")


def _get_expert_ckpt_name(path, expert_id, iteration, mp_rank=0, release=False) -> str:
    return os.path.join(_iter_dir(path, iteration, release), f"expert_{expert_id}_mp_rank_{mp_rank:02d}_model_states.pt")


def _get_optimizer_ckpt_name(path, iteration, expert_parallel_rank, mp_rank=0, release=False) -> str:
    return os.path.join(_iter_dir(path, iteration, release), f"expert_parallel_rank_{expert_parallel_rank}_mp_rank_{mp_rank:02d}_optim_states.pt")


def _get_checkpoint_tracker_filename(path: str) -> str:
    return os.path.join(path, "latest_checkpointed_iteration.txt")


def _ensure_directory_exists(filename: str):
    os.makedirs(os.path.dirname(filename), exist_ok=True)


def _read_metadata(tracker_filename: str) -> Tuple[int, bool]:
    with open(tracker_filename, "r") as f:
        meta = f.read().strip()
    try:
        return int(meta), False
    except ValueError:
        if meta != "release":
            logger.error("Invalid metadata file %s. Exiting", tracker_filename)
            sys.exit()
        return 0, True


def _has_moe_layers(model: torch.nn.Module) -> Tuple[bool, int]:
    from ..parallel.moe.layer import MoE

    has, num = False, 0
    for m in model.modules():
        if isinstance(m, MoE):
            has, num = True, m.num_experts
            break
    return has, num


def _rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def _world() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def _barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def _split_moe_state_dict(full: Dict[str, torch.Tensor], num_local_experts: int, ep_rank: int):
    """(per-global-expert state dicts, non-expert state dict); expert keys are renamed local id → global id."""
    experts: Dict[str, Dict[str, torch.Tensor]] = defaultdict(dict)
    rest: Dict[str, torch.Tensor] = {}
    pat = re.compile(f".*{re.escape(_MOE_PREFIX)}([0-9]+).*")
    for key, value in full.items():
        if "expert" in key and "moe.gate.wg.weight" not in key:
            m = pat.match(key)
            if not m:
                logger.warning("No expert found in key %s.", key)
                rest[key] = value
                continue
            local_id = int(m.group(1))
            gid = ep_rank * num_local_experts + local_id
            experts[str(gid)][key.replace(f"{_MOE_PREFIX}{local_id}", f"{_MOE_PREFIX}{gid}")] = value
        else:
            rest[key] = value
    return experts, rest


def save_checkpoint(iteration: int, checkpoints_path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None,
                    lr_scheduler=None):
    """Save model (and optimizer / lr scheduler) state of ``iteration`` under ``checkpoints_path``.

    Collective: every rank must call it (expert ranks write their experts; rank 0 writes the shared part and the tracker)."""
    logger.info("saving checkpoint at iteration %7d to %s", iteration, checkpoints_path)
    has_moe, num_experts = _has_moe_layers(model)
    # optimizers whose state is sharded across ranks (in-bucket fused SGD/Adam) consolidate it with a collective: every rank
    # has to take part even though only rank 0 writes the result
    collective_opt_sd = optimizer.state_dict() if (optimizer is not None and getattr(optimizer, "collective_state_dict", False)) else None
    if has_moe:
        ep_rank = _rank()
        num_local = num_experts // _world()
        experts_sd, model_sd = _split_moe_state_dict(model.state_dict(), num_local, ep_rank)
        for gid, sd in experts_sd.items():
            name = _get_expert_ckpt_name(checkpoints_path, gid, iteration)
            _ensure_directory_exists(name)
            torch.save(sd, name)
        opt_name = _get_optimizer_ckpt_name(checkpoints_path, iteration, ep_rank)
        _ensure_directory_exists(opt_name)
        torch.save({"optimizer": (collective_opt_sd if collective_opt_sd is not None else optimizer.state_dict()) if optimizer else None}, opt_name)
        if ep_rank == 0:
            state = {"iteration": iteration, "model": model_sd}
            if lr_scheduler is not None:
                state["lr_scheduler"] = lr_scheduler.state_dict()
            name = _get_model_ckpt_name(checkpoints_path, iteration)
            _ensure_directory_exists(name)
            torch.save(state, name)
    elif _rank() == 0:
        state = {"iteration": iteration, "model": model.state_dict()}
        if optimizer is not None:
            state["optimizer"] = collective_opt_sd if collective_opt_sd is not None else optimizer.state_dict()
        if lr_scheduler is not None:
            state["lr_scheduler"] = lr_scheduler.state_dict()
        name = _get_model_ckpt_name(checkpoints_path, iteration)
        _ensure_directory_exists(name)
        torch.save(state, name)
    _barrier()
    if _rank() == 0:
        os.makedirs(checkpoints_path, exist_ok=True)
        with open(_get_checkpoint_tracker_filename(checkpoints_path), "w") as f:
            f.write(str(iteration))
    _barrier()
    logger.info("successfully saved checkpoint at iteration %7d", iteration)


def load_checkpoint(checkpoints_path: str, model: torch.nn.Module, optimizer: Optional[torch.optim.Optimizer] = None, lr_scheduler=None,
                    strict: bool = True) -> int:
    """Load the latest checkpoint under ``checkpoints_path`` into ``model`` (and optimizer / scheduler); returns its
    iteration (0 when there is no checkpoint)."""
    tracker = _get_checkpoint_tracker_filename(checkpoints_path)
    if not os.path.isfile(tracker):
        logger.warning("could not find checkpoint metadata file %s, will not load any checkpoint", tracker)
        return 0
    iteration, release = _read_metadata(tracker)
    logger.info("loading checkpoint at iteration %d from %s", iteration, checkpoints_path)
    ep_rank = _rank()
    ckpt = torch.load(_get_model_ckpt_name(checkpoints_path, iteration, release=release), map_location="cpu", weights_only=False)
    has_moe, num_experts = _has_moe_layers(model)
    if has_moe:
        num_local = num_experts // _world()
        for local_id in range(num_local):
            gid = ep_rank * num_local + local_id
            esd = torch.load(_get_expert_ckpt_name(checkpoints_path, str(gid), iteration, release=release), map_location="cpu", weights_only=False)
            for key in list(esd.keys()):
                ckpt["model"][key.replace(f"{_MOE_PREFIX}{gid}", f"{_MOE_PREFIX}{local_id}")] = esd.pop(key)
    if has_moe and optimizer is not None:
        optim_ckpt = torch.load(_get_optimizer_ckpt_name(checkpoints_path, iteration, ep_rank, release=release), map_location="cpu", weights_only=False)
    else:
        optim_ckpt = ckpt
    model.load_state_dict(ckpt["model"], strict=strict)
    if optimizer is not None and optim_ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(optim_ckpt["optimizer"])
    elif optimizer is not None and hasattr(optimizer, "refresh_master_weights"):
        optimizer.refresh_master_weights()  # weights changed under an optimizer that keeps fp32 master copies
    if lr_scheduler is not None and "lr_scheduler" in ckpt:
        lr_scheduler.load_state_dict(ckpt["lr_scheduler"])
    _barrier()
    logger.info("successfully loaded checkpoint at iteration %d", iteration)
    return iteration
