"""Experiment: the C3 log-lik batch (N=512 x 4096, ARD D=4) split over L concurrent streams ("lanes"), each lane working
through its share in sub-batches, so that one lane's latency-bound panel factor launches can overlap another lane's
DMMA launches.  Prints ms per full pass for several (lanes, sub-batch) combinations."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmc_b200 as gp
import torch

n, B, D = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (512, 4096, 4)
if D > 1:
    x, _ = gp.synthetic.ard_inputs(n, D)
else:
    x = np.arange(n, dtype=np.float64).reshape(n, 1)
G, H = gp.synthetic.loglik_batch(B, n, n_ell=D if D > 1 else 1)
xd, Gd, Hd = torch.tensor(x).cuda(), torch.tensor(G).cuda(), torch.tensor(H).cuda()
ref, _ = gp.ops.loglik_batched(xd, Gd, Hd)
for lanes, sub in ((1, B), (2, B // 2), (2, B // 4), (4, B // 4), (4, B // 8), (4, B // 16), (8, B // 8), (8, B // 16), (8, B // 32)):
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    wss = [gp.ops.Workspace() for _ in range(lanes)]
    chunks = [(s0, min(B, s0 + sub)) for s0 in range(0, B, sub)]
    outs = [None] * len(chunks)

    def run():
        torch.cuda.synchronize()
        for k, (a, b) in enumerate(chunks):
            with torch.cuda.stream(streams[k % lanes]):
                outs[k] = gp.ops.loglik_batched(xd, Gd[a:b], Hd[a:b], workspace=wss[k % lanes], jitter_policy=gp.JITTER_NONE)[0]
        torch.cuda.synchronize()
    for _ in range(2):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 4
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    got = torch.cat(outs)
    print(json.dumps({'n': n, 'B': B, 'lanes': lanes, 'sub_batch': sub, 'ms_per_pass': round(e0.elapsed_time(e1) / reps, 3),
                      'bit_equal': bool(torch.equal(got, ref))}))
