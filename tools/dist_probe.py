"""Timing probe for dist_kernel (kernel tuning only): varies set size, query tile, outputs and the p-value epilogue."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mash_b200
from bench import make_sketches_device

dev = torch.device("cuda", 0)
eng = mash_b200.Engine(0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
S, K = 1000, 21
ks = 4.0 ** K


def run(n, q_count, outs="ndpvx", max_distance=1.0, fam=100, reps=2):
    H, N, L = make_sketches_device(torch, dev, n, S, seed=5, n_families=fam)
    ref = mash_b200._capi._Set(H.data_ptr(), N.data_ptr(), L.data_ptr(), on_device=True, n=n, stride=S)
    job = mash_b200._capi.DistJob(eng, ref, None, None, None, None, None, S, K, ks, max_distance, 1.0)
    q_count = min(q_count, n)
    np_ = q_count * n
    bufs = {"n": torch.empty(np_, dtype=torch.int32, device=dev) if "n" in outs else None,
            "d": torch.empty(np_, dtype=torch.int32, device=dev) if "d" in outs else None,
            "p": torch.empty(np_, dtype=torch.float64, device=dev) if "p" in outs else None,
            "v": torch.empty(np_, dtype=torch.float64, device=dev) if "v" in outs else None,
            "x": torch.empty(np_, dtype=torch.uint8, device=dev) if "x" in outs else None}
    ptr = lambda k: bufs[k].data_ptr() if bufs[k] is not None else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    job.run_dev(0, q_count, ptr("n"), ptr("d"), ptr("p"), ptr("v"), ptr("x"), stream=st.cuda_stream)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        job.run_dev(0, q_count, ptr("n"), ptr("d"), ptr("p"), ptr("v"), ptr("x"), stream=st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"n={n} q={q_count} outs={outs} maxdist={max_distance} fam={fam}: {ms:.2f} ms  {np_ / ms / 1e6:.2f} Gpairs/s", flush=True)
    job.close()
    del bufs, H, N, L
    torch.cuda.empty_cache()


run(16384, 16384)
run(16384, 16384, outs="n")
run(16384, 16384, max_distance=1e-9)
run(16384, 16384, fam=16384)
run(100000, 5368)
run(100000, 5368, outs="n")
run(100000, 5368, max_distance=1e-9)
run(100000, 5368, fam=100000)
run(100000, 1342)
run(8192, 8192)
