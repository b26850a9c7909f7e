"""Times g4c_edge_aggr_fwd alone on a synthetic fixed-k topology (scratch benchmark, not the driver's bench.py).
    python tools/bench_edge.py [--nodes N] [--k K] [--reps R] [--spread S]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphs4cfd_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=1_000_000)
ap.add_argument("--k", type=int, default=6)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--spread", type=int, default=2000, help="sources are drawn within +-spread of the target id")
ap.add_argument("--layers", type=int, default=3)
ap.add_argument("--no-e", action="store_true")
ap.add_argument("--act", default="selu")
ap.add_argument("--variants", default="v5", help="comma-separated kernels to time (v3 = csrc/mp_edge_pair.cu, v5 = csrc/mp_edge_v5.cu); "
                                                "the outputs of every variant are compared with the first one's")
a = ap.parse_args()

dev = torch.device("cuda")
torch.manual_seed(0)
n, k = a.nodes, a.k
col = torch.arange(n, device=dev).repeat_interleave(k)
row = (col + torch.randint(-a.spread, a.spread + 1, (n * k,), device=dev)).clamp_(0, n - 1)
dims = [384] + [128] * a.layers
lin = [((torch.rand(dims[i + 1], dims[i], device=dev) * 2 - 1) / dims[i] ** 0.5,
        (torch.rand(dims[i + 1], device=dev) * 2 - 1) / dims[i] ** 0.5) for i in range(a.layers)]
ln = (torch.ones(128, device=dev), torch.zeros(128, device=dev))
pack = ops.EdgePairPack(lin, ln)
e = torch.randn(n * k, 128, device=dev)
P_r = torch.randn(n, 128, device=dev)
P_c = torch.randn(n, 128, device=dev)
topo = ops.MpTopo.from_edge_index(torch.stack([row, col]), n)
e_out = torch.empty_like(e)
agg = torch.empty(n, 128, device=dev)


def launch(variant="auto"):
    ops.edge_aggr(pack, topo, e, P_r, P_c, act_e=(None if a.act == "none" else a.act), want_e=not a.no_e, e_out=e_out, agg_out=agg,
                  variant=variant)


E = n * k
alg = 4 * 128 * ((1 if a.no_e else 2) * E + 3 * n) + 4 * E
flop = 2 * E * 128 * 128 * a.layers
first = None
for variant in a.variants.split(","):
    e_out.fill_(float("nan"))
    agg.fill_(float("nan"))
    for _ in range(3):
        launch(variant)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.reps):
        launch(variant)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.reps
    same = ""
    if first is None:
        first = (agg.clone(), None if a.no_e else e_out.clone())
    else:
        same = f"; vs {a.variants.split(',')[0]}: agg rel-L2 {float((agg - first[0]).norm() / first[0].norm()):.2e}"
        if not a.no_e:
            same += f", e' rel-L2 {float((e_out - first[1]).norm() / first[1].norm()):.2e}"
    print(f"edge kernel {variant}: N={n} E={E} layers={a.layers} write_e={not a.no_e}: {ms:.3f} ms/launch, "
          f"{alg / ms / 1e6:.1f} GB/s algorithmic ({alg / 1e9:.2f} GB), {flop / ms / 1e9:.1f} TFLOP/s useful "
          f"({3 * flop / ms / 1e9:.1f} issued fp16){same}")

if os.environ.get("G4C_PROFILE"):
    import ctypes as C
    import numpy as np
    from graphs4cfd_b200 import _lib as L
    buf = np.zeros(64, dtype=np.uint64)
    variant = a.variants.split(",")[-1]                              # the last variant is the one profiled
    vcode = ops.EDGE_VARIANTS[variant]
    L.lib().g4c_debug_profile(vcode, buf.ctypes.data_as(C.c_void_p))        # drop warm-up + timing launches
    launch(variant)
    L.check(L.lib().g4c_debug_profile(vcode, buf.ctypes.data_as(C.c_void_p)))
    print(f"phase profile of {variant} (laps of lane 0 of one warp per role in CTA 0):")
    names = {0: ("epilogue warp 0", ["wait MMA (hidden)", "wait MMA (last)", "hidden epilogue", "last: statistics", "last: barrier", "last: normalise+agg+store", "unit end", "-"]),
             8: ("loader warp 16", ["wait rows", "wait acc release", "process + prefetch", "-", "-", "-", "-", "-"]),
             16: ("MMA issuer", ["wait loaders", "wait epilogue", "issue", "-", "-", "-", "-", "-"])}
    n_slots = (n + 127) // 128 * k / 148.0
    for base, (who, labels) in names.items():
        vals = buf[base:base + 8].astype(np.float64)
        tot = vals.sum()
        if tot == 0:
            continue
        print(f"{who}: total {tot / 1e6:.3f} Mcycles = {tot / n_slots:.0f} per slot; " +
              ", ".join(f"{l} {100 * v / tot:.1f}% ({v / n_slots:.0f}/slot)" for l, v in zip(labels, vals) if l != "-"))
