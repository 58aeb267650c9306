"""Timing probe for the dist tile prefilter: configs[2] sketches (family-contiguous and shuffled), prefilter off / on,
one query tile of the 100 000 x 100 000 grid.  Kernel tuning only -- bench.py is the measurement of record."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mash_b200
from bench import make_sketches_device

dev = torch.device("cuda", 0)
eng = mash_b200.Engine(0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
S, K = 1000, 21
ks = 4.0 ** K
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
q_count = int(sys.argv[2]) if len(sys.argv) > 2 else 5368

H, N, L = make_sketches_device(torch, dev, n, S, seed=1000)
perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
np_ = q_count * n
o_n = torch.empty(np_, dtype=torch.int32, device=dev); o_d = torch.empty(np_, dtype=torch.int32, device=dev)
o_D = torch.empty(np_, dtype=torch.float64, device=dev); o_p = torch.empty(np_, dtype=torch.float64, device=dev)
o_x = torch.empty(np_, dtype=torch.uint8, device=dev)
results = {}
for order in os.environ.get("ORDERS", "family_contiguous,shuffled").split(","):
    Hs, Ls = (H, L) if order == "family_contiguous" else (H[perm].contiguous(), L[perm].contiguous())
    ref = mash_b200._capi._Set(Hs.data_ptr(), N.data_ptr(), Ls.data_ptr(), on_device=True, n=n, stride=S)
    job = mash_b200._capi.DistJob(eng, ref, None, None, None, None, None, S, K, ks, 1.0, 1.0)
    keep = {}
    for mode in (0, 1):
        job.set_prefilter(mode)
        args = (o_n.data_ptr(), o_d.data_ptr(), o_D.data_ptr(), o_p.data_ptr(), o_x.data_ptr())
        job.run_dev(0, q_count, *args, stream=st.cuda_stream)
        torch.cuda.synchronize()
        eng.set_timing(True); eng.stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        reps = 2
        for _ in range(reps):
            job.run_dev(0, q_count, *args, stream=st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        stt = eng.stats(reset=True); eng.set_timing(False)
        keep[mode] = (o_n.clone(), o_d.clone(), o_D.clone(), o_p.clone(), o_x.clone())
        results[f"{order}/prefilter={mode}"] = {"ms": ms, "Gpairs_per_s": np_ / ms / 1e6, "dist_kernel_ms": stt["dist_kernel_ms"] / reps,
                                               "stats": job.prefilter_stats() if mode else None}
        print(order, mode, results[f"{order}/prefilter={mode}"], flush=True)
    same = all(torch.equal(a, b) for a, b in zip(keep[0], keep[1]))
    results[f"{order}/identical"] = bool(same)
    print(order, "identical outputs:", same, "pairs with shared hashes:", int((keep[0][0] > 0).sum().item()), flush=True)
    del keep
    job.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open(os.environ.get("PROBE_OUT", "gpurun_out/dist_prefilter_probe.json"), "w"), indent=1)
