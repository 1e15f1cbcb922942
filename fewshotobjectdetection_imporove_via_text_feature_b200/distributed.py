"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL on the GPU box, gloo in CPU tests).

The ROI-head path shards by image: every stage is per-image, weights and text embeddings are
replicated, so inference needs NO data-path collective — only the final exchange of detections, which replaces
the reference's pickle-based `comm.gather` (pascal_voc_evaluation.py:84) with two fixed-shape all_gathers.
Fine-tuning adds the gradient all-reduce (reference: DDP at engine/defaults.py:252-258).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_images(num_images, rank=None, world_size=None):
    """Indices owned by `rank`: r, r+W, r+2W, ...  (InferenceSampler semantics, dataloader/build.py:411-421)."""
    if rank is None:
        rank, world_size = world()
    return list(range(rank, num_images, world_size))


def pack_detections(instances_list, max_dets=100):
    """list[Instances] -> (counts (n,) int32, dets (n,max_dets,6) fp32 [x1,y1,x2,y2,score,class])."""
    n = len(instances_list)
    dev = instances_list[0].scores.device if n else torch.device("cpu")
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    dets = torch.zeros((n, max_dets, 6), dtype=torch.float32, device=dev)
    for i, inst in enumerate(instances_list):
        k = min(len(inst), max_dets)
        counts[i] = k
        if k:
            dets[i, :k, :4] = inst.pred_boxes.tensor[:k]
            dets[i, :k, 4] = inst.scores[:k]
            dets[i, :k, 5] = inst.pred_classes[:k].float()
    return counts, dets


def all_gather_detections(counts, dets, num_images_total):
    """Exchange per-rank detections; returns (counts (num_images_total,), dets (num_images_total,max,6)) ordered by
    global image index (rank r owns images r::W).  Ranks may own different numbers of images: pad to the max."""
    rank, W = world()
    if W == 1:
        return counts, dets
    per = (num_images_total + W - 1) // W
    pc = counts.new_zeros(per)
    pd = dets.new_zeros((per,) + tuple(dets.shape[1:]))
    pc[: counts.shape[0]] = counts
    pd[: dets.shape[0]] = dets
    gc = [torch.empty_like(pc) for _ in range(W)]
    gd = [torch.empty_like(pd) for _ in range(W)]
    dist.all_gather(gc, pc)
    dist.all_gather(gd, pd)
    out_c = counts.new_zeros(num_images_total)
    out_d = dets.new_zeros((num_images_total,) + tuple(dets.shape[1:]))
    for r in range(W):
        idx = list(range(r, num_images_total, W))
        out_c[idx] = gc[r][: len(idx)]
        out_d[idx] = gd[r][: len(idx)]
    return out_c, out_d


def allreduce_gradients(params, bucket_bytes=64 << 20, average=True):
    """Bucketed gradient all-reduce (flat fp32 buckets; NVSwitch makes cost launch-latency-, not link-bound, so
    buckets are sized for overlap rather than link count).  Returns the number of buckets reduced."""
    rank, W = world()
    # buckets are laid out over the FIXED parameter list (a rank whose batch left a parameter without a gradient still
    # contributes zeros), so every rank issues the same collectives with the same sizes
    params = [p for p in params if p.requires_grad]
    if W == 1 or not params:
        return 0
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    grads = [p.grad for p in params]
    buckets, cur, size = [], [], 0
    for g in grads:
        nb = g.numel() * g.element_size()
        if cur and size + nb > bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(g)
        size += nb
    if cur:
        buckets.append(cur)
    works = []
    for b in buckets:
        flat = torch.cat([g.reshape(-1).float() for g in b])
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), flat, b))
    for w, flat, b in works:
        w.wait()
        if average:
            flat.div_(W)
        off = 0
        for g in b:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
    return len(buckets)
