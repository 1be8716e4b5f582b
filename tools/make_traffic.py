#!/usr/bin/env python
"""Writes profiles/traffic.json from `ncu --page raw --csv` dumps of ONE step of bench.py (tools/r2_final.sh).

    python tools/make_traffic.py <source_sha> <capture note> noise:K10:N32=raw1.csv natural:K2:N32=raw2.csv ...

Per workload and pipeline stage (kernels mapped onto bench.py's stage names; the mean over the launches of a stage
in the captured step): DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), L2 sectors, L2 atomic / reduction
requests (compare-and-swap / ALU atomics, reductions) and their sectors at the L1, the L2's own sector count with its
percentage of the hardware's sustained sector rate (the random-sector rate is the bound profiles/README argues for on
dense lattices), warp-level global-memory requests, launch time under ncu.  `source_sha` is bench.source_sha() of the build
that was captured: bench.py drops the file when the sources have changed since.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
STAGE = (("prepare_kernel", "prepare"), ("build_dedup_kernel", "build"), ("build_kernel", "build"),
         ("neighbour_kernel", "neighbour"), ("vertex_init_kernel", "splat"), ("splat_rows_kernel", "splat"),
         ("splat_kernel", "splat"), ("blur_kernel", "blur"), ("slice_kernel", "slice"),
         ("loss_backward_logits_kernel", "backward"), ("loss_backward_kernel", "backward"))
FIELDS = ("dram_bytes", "lts_sectors", "lts_sectors_pct_of_peak", "lts_requests", "lts_requests_atom", "lts_requests_red",
          "l1_sectors_atom", "l1_sectors_red", "l1_global_requests", "ncu_us")
REQ = ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
       "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum")


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        if name not in idx or r[idx[name]] in ("", "no data", "n/a"):
            return None
        return float(r[idx[name]].replace(",", "")) * SCALE.get(units[idx[name]], 1.0)

    acc = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        stage = next((s for k, s in STAGE if k in name), None)
        if stage is None:
            continue
        a = acc.setdefault(stage, dict({k: 0.0 for k in FIELDS}, launches=0, kernels=set()))
        a["launches"] += 1
        a["kernels"].add(name.split("(")[0].replace("void ", ""))
        a["dram_bytes"] += (val(r, "dram__bytes_read.sum") or 0.0) + (val(r, "dram__bytes_write.sum") or 0.0)
        a["lts_sectors"] += val(r, "lts__t_sectors.sum") or 0.0
        a["lts_sectors_pct_of_peak"] += val(r, "lts__t_sectors.sum.pct_of_peak_sustained_elapsed") or 0.0
        a["lts_requests"] += val(r, "lts__t_requests.sum") or 0.0
        a["lts_requests_atom"] += (val(r, "lts__t_requests_srcunit_tex_op_atom_dot_cas.sum") or 0.0) + \
                                  (val(r, "lts__t_requests_srcunit_tex_op_atom_dot_alu.sum") or 0.0)
        a["lts_requests_red"] += val(r, "lts__t_requests_srcunit_tex_op_red.sum") or 0.0
        a["l1_sectors_atom"] += val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum") or 0.0
        a["l1_sectors_red"] += val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum") or 0.0
        a["l1_global_requests"] += sum(val(r, m) or 0.0 for m in REQ)
        a["ncu_us"] += val(r, "gpu__time_duration.sum") or 0.0
    out = {}
    for stage, a in acc.items():
        n = a["launches"]
        out[stage] = {k: a[k] / n for k in FIELDS}
        out[stage]["launches_in_capture"] = n
        out[stage]["kernels"] = sorted(a["kernels"])
    return out


def main():
    sha, note = sys.argv[1], sys.argv[2]
    doc = {"_meta": {"source_sha": sha, "capture": note,
                     "what": "per launch, mean over the launches of a stage in one captured step; written by "
                             "tools/make_traffic.py from ncu --set full raw pages (profiles/*_ncu_full_raw_*.csv)"},
           "workloads": {}}
    for spec in sys.argv[3:]:
        key, path = spec.split("=", 1)
        doc["workloads"][key] = parse(path)
        doc["_meta"].setdefault("files", {})[key] = os.path.basename(path)
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    for key, w in doc["workloads"].items():
        print(key, {s: round(v["dram_bytes"] / 1e6, 1) for s, v in w.items()})


if __name__ == "__main__":
    main()
