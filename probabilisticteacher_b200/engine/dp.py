"""Data-parallel plumbing (one process per GPU, torch.distributed): parameter broadcast at start-up
(`pt/engine/trainer.py:491-496`, DDP `_sync_params_and_buffers`) and the per-step gradient all-reduce
that torch DDP performs for the reference (`trainer.py:92-95,384`): SUM all-reduce over the flat fp32 gradient
arena, with the step's loss scalars riding in a 16-float tail of the same buffer (the reference's cross-rank metric
mean, `trainer.py:394-429`). The 1/world averaging is folded into the clip/SGD kernel (`pre_scale`).
Eager steps issue ONE collective after the backward (`allreduce_grads`; `bucket_elems` can cut it into several). The
CUDA-graph step of `engine/trainer.py` (`_graph_body_concurrent`) issues two buckets inside the graph: everything
behind the VGG backbone as soon as both student passes have finished their head backward (overlapping the backbone
backward), the backbone bucket after the join; the metrics tail follows eagerly. A graph that holds captured NCCL
kernels must be released before `dist.destroy_process_group()` (`PTrainer.release_graphs`)."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_params(flat_params, src=0):
    if world_size() > 1:
        dist.broadcast(flat_params, src=src)
    return flat_params


def bucket_bounds(numel, bucket_elems):
    """Splits [0, numel) into contiguous buckets of at most bucket_elems (last one may be short)."""
    out = []
    s = 0
    while s < numel:
        e = min(numel, s + bucket_elems)
        out.append((s, e))
        s = e
    return out


def allreduce_grads(flat_grads, bucket_elems=32 * 1024 * 1024, async_op=False):
    """SUM all-reduce of the flat gradient arena in buckets; returns the list of work handles (empty
    when world == 1). The caller applies pre_scale = 1 / world when consuming the gradients."""
    if world_size() == 1:
        return []
    works = []
    for s, e in bucket_bounds(flat_grads.numel(), bucket_elems):
        w = dist.all_reduce(flat_grads[s:e], op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def pre_scale():
    return 1.0 / world_size()
