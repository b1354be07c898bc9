"""DRAM traffic per launch of the dominant kernel (omni::gemm_bf16_tn_2cta, plain epilogue) from an ncu capture of ONE train
step, written where bench.py reads it:

   ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \\
       --clock-control none -k regex:gemm_bf16_tn_2cta --csv --log-file gpurun_out/gemm_traffic.csv \\
       python tools/one_step.py --batch 32
   python tools/ncu_traffic.py gpurun_out/gemm_traffic.csv profiles/gemm_traffic.json

The JSON carries the digest of csrc/ + include/ (omni_avsr_b200.build._digest) of the code that was profiled; bench.py
reports `traffic` from it and says whether the digest still matches the library it runs."""
import csv
import json
import sys
from collections import defaultdict


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if r and r[0] == "ID")
    body = rows[rows.index(hdr) + 1:]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = defaultdict(dict)
    for r in body:
        if len(r) <= vi:
            continue
        try:
            per[(r[ii], r[ki])][r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    plain = [(k, m) for k, m in per.items() if "gemm_bf16_tn_2cta<5, 0>" in k[1] or "gemm_bf16_tn_2cta<5,0>" in k[1].replace(" ", "")]
    allp = [(k, m) for k, m in per.items() if "gemm_bf16_tn_2cta" in k[1]]

    def agg(items):
        n = len(items)
        rd = sum(m.get("dram__bytes_read.sum", 0.0) for _, m in items)
        wr = sum(m.get("dram__bytes_write.sum", 0.0) for _, m in items)
        t = sum(m.get("gpu__time_duration.sum", 0.0) for _, m in items)
        return {"launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "time_ns": t,
                "mean_bytes_per_launch": (rd + wr) / max(n, 1)}
    sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
    from omni_avsr_b200 import build as b
    dig = b._digest(sorted(b.CSRC.glob("*.cu")) + sorted(b.CSRC.glob("*.cuh")) + sorted((b.ROOT / "include").glob("*.h")))
    out = {"kernel": "omni::gemm_bf16_tn_2cta<5, 0> (plain epilogue)", "source_digest": dig, "plain": agg(plain),
           "all_pair_kernel_variants": agg(allp),
           "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over "
                  "every pair-kernel launch of one B=32 train step (tools/one_step.py); cold-cache, serialised"}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
