"""profiles/<round>_summary.json from the round's `ncu --set full` captures: per-launch DRAM traffic (dram__bytes_read.sum +
dram__bytes_write.sum) of the kernels bench.py quotes in its `traffic` fields.   usage: python tools/make_summary.py r2s r2"""
import csv
import io
import json
import subprocess
import sys

SRC, DST = sys.argv[1], sys.argv[2]
SCALE = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}


def rows(name):
    out = subprocess.run(["ncu", "-i", f"gpurun_out/{SRC}_{name}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    h, u = r[0], r[1]
    res = []
    for x in r[2:]:
        d = {"name": x[h.index("Kernel Name")]}
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"):
            i = h.index(k)
            d[k] = float(x[i]) * SCALE.get(u[i], 1)
        res.append(d)
    return res


def traffic(d):
    return d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]


S = {"_comment": "per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) from this round's `ncu --set full --clock-control none` "
                 "captures (tools/run_profiles.sh; condensed in profiles/%s_ncu_*.txt). bench.py copies dram_bytes_per_launch into its `traffic` fields." % DST}
g_small, g_vae = rows("gemm"), rows("gemm_vae")
tot = [traffic(x) for x in g_small + g_vae]
S["k_tc_gemm"] = {"dram_bytes_per_launch": int(sum(tot) / len(tot)), "launches": len(tot), "min_max": [int(min(tot)), int(max(tot))],
                  "tensor_pipe_pct_elapsed_vae": [round(x["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"], 1) for x in g_vae],
                  "note": "mean of %d consecutive U-Net launches (outputs stay in the 126 MB L2 for the next kernel: ~0 written) and %d VAE-encoder launches at 512^2 / 256^2 "
                          "(the 512^2 x 128-channel activations are 67 MB each: input + residual + output exceed L2 only there); algorithmic operand + output "
                          "bytes of the 512^2 conv3x3 128->128 are 134 MB -> no re-reads" % (len(g_small), len(g_vae)),
                  "source": "profiles/%s_ncu_gemm.txt, profiles/%s_ncu_gemm_vae.txt" % (DST, DST)}
f = rows("field_fwd")[0]
S["k_field_fwd_tc"] = {"dram_bytes_per_launch": int(traffic(f)), "ms": round(f["gpu__time_duration.sum"] * 1e3, 3),
                       "note": "read: the 48.8 MB table once (gathers are L2 hits); written: encoding cache for the backward + tape / sigmas / rgbs; M = 424 k samples",
                       "source": "profiles/%s_ncu_field_fwd.txt" % DST}
b = rows("field_bwd")[0]
S["k_field_bwd_tc"] = {"dram_bytes_per_launch": int(traffic(b)), "ms": round(b["gpu__time_duration.sum"] * 1e3, 3),
                       "tensor_pipe_pct_elapsed": round(b["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"], 1),
                       "note": "regulariser backward with the table-gradient REDs issued from the chain kernel: read = cached encodings + tape; the REDs resolve in L2 "
                               "(the 1.07 GB encoding-gradient round trip of the split pipeline is gone)", "source": "profiles/%s_ncu_field_bwd.txt" % DST}
S["k_bwd_enc_scatter"] = {"dram_bytes_per_launch": None, "note": "not launched on the default (fused-scatter) path; A/B timings of the split pipeline in profiles/%s_field_ab.txt" % DST}
gn = [x for x in rows("gn_attn") if "k_gn" in x["name"]]
if gn:
    S["k_gn_stats|k_gn_apply"] = {"dram_bytes_per_launch": int(sum(traffic(x) for x in gn) / len(gn)),
                                  "note": "U-Net level-0 GroupNorm (fp16 [2,64,64,320] = 5.2 MB): mostly an L2 hit", "source": "profiles/%s_ncu_gn_attn.txt" % DST}
json.dump(S, open(f"profiles/{DST}_summary.json", "w"), indent=2)
print(json.dumps({k: v.get("dram_bytes_per_launch") if isinstance(v, dict) else None for k, v in S.items()}))
