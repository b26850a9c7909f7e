"""Per-slot cycle budget of the fused edge kernel (csrc/mp_edge_pair.cu and the TMA variants of csrc/mp_edge_pair_tma.cu).

A "slot" is the unit of work of one CTA: 128 edges x 128 columns through the 2- or 3-layer MLP, LayerNorm, aggregation.
For every resource of the SM this prints how many cycles one slot needs, next to the HBM-bound slot time.  It is the
arithmetic behind DESIGN.md 4.1 / 4.1b, kept as a script so that the next measurements can be compared with it.
Inputs that are measurements: HBM peak and SM clock (MEASURED_PEAKS.json / the bench's clocks line), 64 cycles per
M=256 N=128 K=16 MMA (tools/mma_rate.py), the instruction counts of the SASS (cuobjdump of the built kernel), the slot time
of the v3 kernel (profiles/r1h_*).  Inputs that are ASSUMPTIONS of the model (marked *): one 128-byte line per cycle
through the load/store pipe for LDGSTS / STG, four wavefronts per conflict-free LDS.128 / STS.128.

    python tools/edge_kernel_model.py [--k 6] [--layers 3] [--sm-mhz 1900]
"""
import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=6)
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--sm-mhz", type=float, default=1900.0)
    ap.add_argument("--measured-ms", type=float, default=2.86, help="v3 launch time at 1M targets, for the measured slot time")
    a = ap.parse_args()
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6451.5
    n_sm, H, rows = 148, 128, 128
    row_b = 4 * H
    # ---- HBM: e in, e' out per edge; P_c, agg per target (1/k per edge); src id; P_r[src] mostly from L2
    dram_per_edge = 2 * row_b + 2 * row_b / a.k + 4
    bytes_per_cycle_sm = hbm * 1e9 / n_sm / (a.sm_mhz * 1e6)
    t_hbm = rows * dram_per_edge / bytes_per_cycle_sm
    # ---- tensor pipe: per layer 8 K-steps x 3 split products; one M=256 MMA serves the two CTAs of the pair
    t_mma = a.layers * 8 * 3 * 64
    # ---- issue slots: SASS instruction counts per element (cuobjdump): hidden epilogue 9 per element,
    #      last epilogue ~ 12, loaders ~ 5 (split of e: 3, add + scale of P: 2); 4 schedulers
    elems = rows * H
    instr = elems * ((a.layers - 1) * 9 + 12 + 5) / 32
    t_issue = instr / 4
    # ---- load/store pipe (*): line passes
    def lsu(mode):
        ldgsts_lines = {0: 3, 1: 3, 2: 1, 3: 0}[mode] * rows * row_b / 64 / 8 * 8      # 64-byte pieces, 8 rows (= 8 lines) per instruction
        lds_convert = 3 * rows * row_b / 16 / 32 * 4                                    # lane = row LDS.128, 4 wavefronts each
        if mode == 0:
            stores = rows * row_b / 32 / 32 * 32                                        # STG.256: 32 rows = 32 lines per instruction
        else:
            stores = rows * row_b / 16 / 32 * 4                                         # STS.128 into the staging tile
        consts = 16 * (a.layers - 1) * 8 + 16 * 24 + 16 * 10                            # broadcast constant loads, LN partials
        return ldgsts_lines + lds_convert + stores + consts
    # ---- TMEM (64 B/cycle read, 256 B/cycle write): epilogue reads of the accumulator, loader / epilogue writes
    t_tmem_rd = a.layers * rows * row_b / 64
    t_tmem_wr = (rows * row_b + rows * H * 2 * 2 * a.layers) / 256
    n_slots = (1_000_000 / 128) * a.k / n_sm
    t_meas = a.measured_ms * 1e-3 * a.sm_mhz * 1e6 / n_slots
    print(f"edge kernel, k = {a.k}, {a.layers} layers, SM clock {a.sm_mhz:.0f} MHz, HBM {hbm:.0f} GB/s = {bytes_per_cycle_sm:.1f} B/cycle/SM")
    print(f"  DRAM bytes per edge {dram_per_edge:.0f}  ->  HBM-bound slot time        {t_hbm:8.0f} cycles   (= 100 % of the roofline)")
    print(f"  tensor pipe ({a.layers * 24} pair MMAs x 64)                      {t_mma:8.0f} cycles   ({100 * t_mma / t_hbm:.0f} % of it)")
    print(f"  issue slots ({instr:.0f} warp instructions / 4 schedulers)    {t_issue:8.0f} cycles   ({100 * t_issue / t_hbm:.0f} %)")
    print(f"  TMEM reads by the epilogue {t_tmem_rd:.0f}, TMEM writes {t_tmem_wr:.0f} cycles")
    for mode, name in ((0, "v3 (default): LDGSTS x3, STG.256"), (1, "mode 1: e' through smem + TMA store"),
                       (2, "mode 2: + e, P_c by TMA tile loads"), (3, "mode 3: + P_r by TMA gather4")):
        t = lsu(mode)
        print(f"  load/store pipe*, {name:40s} {t:8.0f} cycles   ({100 * t / t_hbm:.0f} %)")
    print(f"  measured v3 slot time ({a.measured_ms} ms per launch at 1M targets)      {t_meas:8.0f} cycles   ({100 * t_hbm / t_meas:.0f} % of the roofline)")
    print("  (* model assumption, see the docstring; the in-kernel phase profile of v3, profiles/r1d_edge_pair_v3_phases.txt, shows the epilogue")
    print("   warps busy ~16.6k cycles per slot, 9.5k of them in the store tail)")


if __name__ == "__main__":
    main()
