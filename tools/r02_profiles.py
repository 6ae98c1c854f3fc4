#!/usr/bin/env python
"""Turn the round-2 captures of tools/gpu_r02_final.sh (gpurun_out/r02_*) into the tracked summaries under profiles/.

    python tools/r02_profiles.py            # reads gpurun_out/, writes profiles/r02_*

 * r02_launches_bench.csv / r02_launches_train.csv        per-kernel device time (tools/launch_shares.py of the launch lists)
 * r02_mlp_tc3_ncu_full_key_metrics.csv                   ncu --set full of the forward kernel as the bench launches it
 * r02_mlp_tc3_ncu_bench.json                             the figures bench.py's roofline.traffic is scaled from
 * r02_train_kernels_ncu_full_key_metrics.csv             forward (training mode), dgrad chain, weight-gradient kernels
 * r02_sass_index.txt                                     tcgen05 / TMEM / bulk-copy mnemonics per kernel (cuobjdump -sass)
"""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, ROOT)


def shares(name):
    src = os.path.join(OUT, name)
    if not os.path.exists(src):
        return
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), src], capture_output=True, text=True).stdout
    with open(os.path.join(PROF, name), "w") as f:
        f.write(txt)


def copy_keys(src, dst):
    s = os.path.join(OUT, src)
    if os.path.exists(s):
        with open(s) as f, open(os.path.join(PROF, dst), "w") as g:
            g.write(f.read())


def bench_json():
    src = os.path.join(OUT, "r02_mlp_tc3_bench.csv")
    if not os.path.exists(src):
        return
    m = {r[1]: float(r[3]) for r in list(csv.reader(open(src)))[1:]}
    unit = {r[1]: r[2] for r in list(csv.reader(open(src)))[1:]}

    def to_bytes(key):
        return m[key] * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit[key]]
    import bench
    pixels, poses, s_fine = 65536, bench.N_POSES, bench.S_C + bench.N_I
    rows = pixels * poses * s_fine
    macs = 63 * 256 + 4 * 65536 + 319 * 256 + 2 * 65536 + 256 + 65536 + 283 * 128 + 128 * bench.CH      # the reference's MACs per sample
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    # algorithmic bytes of the fine pass with compositing in the kernel: per sample the depth in (4 B) and the density out (4 B, the
    # `sigma` output Graph.render returns); per ray origin + direction (24 B) and the view bias (128 fp32) in, rgb / disp / acc out
    alg = rows * (4 + 4) + pixels * poses * (24 + 128 * 4 + 4 * (bench.CH + 2))
    doc = {"kernel": "bnrf::tc3::mlp_tc3_kernel<3,false,true> (4th MLP launch of a bench step: fine pass of the 19-pose blur render, compositing fused)",
           "rows": rows, "algorithmic_flop": 2.0 * macs * rows, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
           "gpu_time_ms_under_ncu": m["gpu__time_duration.sum"], "sm_ghz_under_ncu": m["sm__cycles_elapsed.max.per_second"],
           "tensor_pipe_active_pct": m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"],
           "l2_hit_rate_pct": m["lts__t_sector_hit_rate.pct"], "algorithmic_bytes": alg,
           "command": "ncu --set full --clock-control none --import-source on -k regex:mlp_tc3_kernel -s 3 -c 1 python bench.py "
                      "--steps 1 --warmup 1 --pixels 65536 --no-cpu-baseline --no-train-step",
           "source": "profiles/r02_mlp_tc3_ncu_full_key_metrics.csv"}
    with open(os.path.join(PROF, "r02_mlp_tc3_ncu_bench.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("tensor pipe %.1f %%, DRAM %.2f GB = %.2f x algorithmic" % (doc["tensor_pipe_active_pct"], (rd + wr) / 1e9, (rd + wr) / alg))


def sass_index():
    """Mnemonic counts per kernel from the built objects: the evidence that the hot kernels are tcgen05 / TMEM / bulk-copy code."""
    build = os.path.join(ROOT, "benerf_b200", "build")
    pats = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "ELECT", "HMMA", "LDL", "STL"]
    lines = ["# cuobjdump -sass of benerf_b200/build/*.o (sm_100a): instruction counts per kernel", "# kernel | " + " | ".join(pats)]
    for obj in sorted(os.listdir(build)):
        if not obj.endswith(".o"):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(build, obj)], capture_output=True, text=True).stdout
        cur, counts = None, {}
        for ln in sass.splitlines():
            mm = re.search(r"Function : (\S+)", ln)
            if mm:
                cur = subprocess.run(["c++filt", mm.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
                counts[cur] = dict.fromkeys(pats, 0)
                continue
            if cur:
                for p in pats:
                    if re.search(r"\b" + re.escape(p) + (r"\b" if "." not in p else ""), ln):
                        counts[cur][p] += 1
        for k, c in counts.items():
            if c["UTCHMMA"] or c["UBLKCP"] or c["LDTM"]:
                lines.append("%s (%s) | " % (k, obj) + " | ".join(str(c[p]) for p in pats))
    with open(os.path.join(PROF, "r02_sass_index.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    shares("r02_launches_bench.csv")
    shares("r02_launches_train.csv")
    copy_keys("r02_mlp_tc3_bench.csv", "r02_mlp_tc3_ncu_full_key_metrics.csv")
    copy_keys("r02_train_kernels.csv", "r02_train_kernels_ncu_full_key_metrics.csv")
    for name in ("r02_mlp_tc3_stall_trace.txt", "r02_bwd_stall_trace.txt", "r02_parity_report.json", "r02_bench_n1.json", "r02_bench_n2.json",
                 "r02_bench_n8.json", "r02_bench_reference.json"):
        copy_keys(name, name)
    copy_keys("gradient_parity.json", "r02_gradient_parity.json")
    copy_keys("parity_bench_batch_sample.json", "r02_parity_bench_batch_sample.json")
    bench_json()
    sass_index()
