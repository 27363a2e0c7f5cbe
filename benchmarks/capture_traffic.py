#!/usr/bin/env python
"""Turns ONE `ncu --set full --page raw --csv` export of a bench.py step into profiles/dram_traffic.json.

    # on the GPU box (gpurun), after `python bench.py ...` has exited 0 without ncu:
    ncu --set full --clock-control none -k regex:'sci_fwd|cci_|sci_bwd|rbf_|dec_' -c 40 -o gpurun_out/step \
        python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --encounters 131072
    ncu -i gpurun_out/step.ncu-rep --page raw --csv > gpurun_out/step_raw.csv
    python benchmarks/capture_traffic.py gpurun_out/step_raw.csv 131072 profiles/dram_traffic.json

Per kernel of the step (last launch of each): DRAM bytes per launch, executed XU (MUFU) warp instructions, issue /
XU / FMA / DRAM utilisation.  The file records the SHA-256 of the kernel sources it was measured on; bench.py prints
`roofline.traffic` and the executed-ex2 view only when that hash matches the tree it runs from (stale numbers are
refused, not printed).
"""
import csv
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from deep_interpolation_clustering_b200.build import sources_sha256  # noqa: E402

NAMES = {"sci_fwd_kernel": "sci_fwd", "cci_fwd": "cci_fwd", "cci_bwd": "cci_sci_bwd", "sci_bwd_kernel": "sci_bwd",
         "rbf_fwd": "rbf_fwd", "rbf_bwd_kernel": "rbf_bwd", "dec_q": "dec_q", "dec_p_kernel": "dec_p",
         "interp_fused_fwd": "sci_cci_fwd", "interp_fused_bwd": "cci_sci_bwd"}


def main():
    raw, enc, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale_by_unit=True):
        if name not in ix or r[ix[name]] in ("", "n/a"):
            return None
        v = float(r[ix[name]].replace(",", ""))
        if scale_by_unit:
            u = units[ix[name]].lower()
            v *= {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}.get(u, 1.0)
        return v

    res = {}
    for r in data:
        kname = r[ix["Kernel Name"]]
        key = next((v for k, v in NAMES.items() if k in kname), None)
        if key is None:
            continue
        res[key] = {
            "bytes_per_launch": (val(r, "dram__bytes_read.sum") or 0) + (val(r, "dram__bytes_write.sum") or 0),
            "xu_warp_inst": val(r, "sm__inst_executed_pipe_xu.sum", False),
            "warp_inst": val(r, "smsp__inst_executed.sum", False),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
            "xu_pipe_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", False),
            "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False),
            "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
            "shared_bank_conflict_share": (val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False) or 0) /
                                          max(val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False) or 1, 1),
            "ncu_duration_us": val(r, "gpu__time_duration.sum", False),
        }
    doc = {"_source": "ncu --set full --clock-control none of one bench.py step (benchmarks/capture_traffic.py); per launch",
           "encounters": enc, "sources_sha256": sources_sha256(), "kernels": res}
    json.dump(doc, open(out, "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
