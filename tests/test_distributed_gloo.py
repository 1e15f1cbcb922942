"""N>1 path on CPU: world_size-2 gloo processes exercise image sharding, the fixed-shape detection exchange that
replaces the reference's pickle gather, and the bucketed gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_images):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fewshotobjectdetection_imporove_via_text_feature_b200 import distributed as D
    from fewshotobjectdetection_imporove_via_text_feature_b200.structures import Boxes, Instances
    mine = D.shard_images(n_images)
    assert mine == list(range(rank, n_images, world))
    insts = []
    for gi in mine:                                   # image gi has gi % 4 detections, all tagged with gi
        k = gi % 4
        insts.append(Instances((10, 10), pred_boxes=Boxes(torch.full((k, 4), float(gi))), scores=torch.full((k,), gi / 10.0),
                               pred_classes=torch.full((k,), gi, dtype=torch.int64)))
    counts, dets = D.pack_detections(insts, max_dets=5) if insts else (torch.zeros(0, dtype=torch.int32), torch.zeros(0, 5, 6))
    gc, gd = D.all_gather_detections(counts, dets, n_images)
    assert gc.tolist() == [i % 4 for i in range(n_images)]
    for i in range(n_images):
        assert bool((gd[i, :i % 4, 5] == i).all()) and bool((gd[i, i % 4:] == 0).all())
    # the evaluators' records after the exchange == the records of one process holding every image
    # (replaces comm.gather of the per-rank `_predictions` dicts, pascal_voc_evaluation.py:84-91)
    from fewshotobjectdetection_imporove_via_text_feature_b200.evaluation import detection_formats as F
    ids = ["%06d" % i for i in range(n_images)]
    lines = F.voc_prediction_lines(ids, gd[..., :4].numpy(), gd[..., 4].numpy(), gd[..., 5].numpy().astype("int64"), gc.numpy())
    want = {}
    for i in range(n_images):
        for _ in range(i % 4):
            want.setdefault(i, []).append(f"{ids[i]} {i / 10.0:.3f} {i + 1:.1f} {i + 1:.1f} {float(i):.1f} {float(i):.1f}")
    assert dict(lines) == want
    # gradient all-reduce: rank r holds grad = r+1 everywhere -> mean 1.5
    ps = [torch.nn.Parameter(torch.zeros(1000)), torch.nn.Parameter(torch.zeros(37, 3)), torch.nn.Parameter(torch.zeros(5))]
    for p in ps[:2]:
        p.grad = torch.full_like(p, float(rank + 1))
    if rank == 1:                      # a parameter that received a gradient on ONE rank only: the bucket layout is over the
        ps[2].grad = torch.full_like(ps[2], 2.0)     # fixed parameter list, the other rank contributes zeros (no hang)
    nb = D.allreduce_gradients(ps, bucket_bytes=2048)
    assert nb >= 2 and all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in ps[:2])
    assert torch.allclose(ps[2].grad, torch.full_like(ps[2], 1.0))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 7), nprocs=2, join=True)


def test_single_process_noop():
    from fewshotobjectdetection_imporove_via_text_feature_b200 import distributed as D
    assert D.world() == (0, 1) and D.shard_images(5) == [0, 1, 2, 3, 4]
    c, d = torch.tensor([1], dtype=torch.int32), torch.zeros(1, 5, 6)
    assert D.all_gather_detections(c, d, 1)[0] is c
