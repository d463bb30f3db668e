"""Bisects run-to-run changes in the proposal path of the small parity fixture: RPN head outputs, key sort, top-k
decode, second sort, NMS -- each stage repeated on FIXED inputs and compared bit for bit."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_parity_x3_gpu as T
from probabilisticteacher_b200 import ops
from probabilisticteacher_b200._lib import call

cuda = torch.device("cuda:0")
H, W, K = 192, 272, 8
O, model, om = T._pair(cuda, K, "DifferentiableAnchorGenerator", 3)
lab = O.synthetic_batch(2, H, W, K, 1)
model.prio_override = T._prios(cuda, 2, H, W, 7)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
with torch.no_grad():
    model(T._to_inst(lab), branch="supervised")
    ctx = model._last_ctx
    feat = ctx["feat"]
    rpn = model.proposal_generator
    # 1. backbone + head repeated
    t0, l0, d0 = rpn.rpn_head(feat)
    l0, d0 = l0.clone(), d0.clone()
    bad = 0
    for i in range(reps):
        t, l, d = rpn.rpn_head(feat)
        bad += int(not (torch.equal(l, l0) and torch.equal(d, d0)))
    print("rpn head changed:", bad, "of", reps, flush=True)
    logits, deltas, anchors = l0, d0, ctx["rpn"]["anchors"]
    N, Hf, Wf, A = 2, feat.H, feat.W, rpn.arena.A
    R = Hf * Wf * A
    k = min(R, rpn.pre_nms_topk[True])
    img_hw = torch.tensor([[H, W]] * 2, dtype=torch.float32, device=cuda)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    first = None
    cnt = {}
    for i in range(reps):
        keys = torch.empty(N, R, dtype=torch.int32, device=cuda)
        vals = torch.empty(N, R, dtype=torch.int32, device=cuda)
        call("ptb200_rpn_make_keys", logits, logits.shape[2], N, Hf, Wf, A, keys, vals)
        ops.segmented_sort(keys, vals)
        boxes = torch.zeros(N, k, 4, dtype=torch.float32, device=cuda)
        scores = torch.zeros(N, k, dtype=torch.float32, device=cuda)
        keys2 = torch.zeros(N, k, dtype=torch.int32, device=cuda)
        vals2 = torch.zeros(N, k, dtype=torch.int32, device=cuda)
        valid = torch.empty(N, dtype=torch.int32, device=cuda)
        call("ptb200_rpn_topk_decode", vals, R, logits, logits.shape[2], deltas, deltas.shape[2], anchors, N, Hf, Wf,
             A, k, img_hw, float(rpn.min_box_size), boxes, scores, keys2, vals2, valid, flag)
        k2u, v2u = keys2.clone(), vals2.clone()
        ops.segmented_sort(keys2, vals2)
        keep_idx, keep_count = ops.nms(boxes, vals2, valid, rpn.nms_thresh, rpn.post_nms_topk[True])
        cur = dict(sort1_keys=keys.clone(), sort1_vals=vals.clone(), boxes=boxes, scores=scores, keys2_unsorted=k2u,
                   vals2_unsorted=v2u, valid=valid.clone(), keys2=keys2.clone(), vals2=vals2.clone(),
                   keep_count=keep_count.clone())
        kc = keep_count.tolist()
        cur["keep_idx"] = torch.stack([torch.where(torch.arange(keep_idx.shape[1], device=cuda) < kc[n], keep_idx[n], -1) for n in range(N)])
        if first is None:
            first = cur
            print("valid", valid.tolist(), "keep", kc, flush=True)
            continue
        for name in cur:
            if not torch.equal(cur[name], first[name]):
                cnt[name] = cnt.get(name, 0) + 1
    print("stages that changed on fixed inputs (count of", reps - 1, "):", cnt, flush=True)
    # NMS alone on the first run's inputs
    badn = 0
    kc0 = None
    for i in range(reps * 3):
        keep_idx, keep_count = ops.nms(first["boxes"], first["vals2"], first["valid"], rpn.nms_thresh, rpn.post_nms_topk[True])
        kc = keep_count.tolist()
        if kc0 is None:
            kc0 = kc
        badn += int(kc != kc0)
    print("nms alone changed:", badn, "of", reps * 3, "first", kc0, flush=True)
