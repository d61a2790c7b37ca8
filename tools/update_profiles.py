#!/usr/bin/env python
"""Copy the evidence of a tools/gpu_profile_final.sh run from gpurun_out/ into profiles/ (tracked).

    python tools/update_profiles.py <tag> [deblock-report-tag] [--round r2]
"""
import json, os, re, shutil, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
RND = "r1"
if "--round" in sys.argv:
    RND = sys.argv[sys.argv.index("--round") + 1]
    argv = [a for a in argv if a != RND]
tag = argv[0]
dbk = argv[1] if len(argv) > 1 else None


def summary(rep):
    return subprocess.run([sys.executable, os.path.join(REPO, "tools", "ncu_summary.py"), rep],
                          capture_output=True, text=True).stdout


def first_json_line(path):
    for line in open(path):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit("no JSON line in " + path)


line = first_json_line(os.path.join(OUT, "bench_%s.json" % tag))
json.dump(line, open(os.path.join(PROF, RND + "_bench_line.json"), "w"), indent=1)
if os.path.exists(os.path.join(OUT, "bench_%s_reference.json" % tag)):
    ref = first_json_line(os.path.join(OUT, "bench_%s_reference.json" % tag))
    json.dump(ref, open(os.path.join(PROF, RND + "_bench_reference_line.json"), "w"), indent=1)
shutil.copy(os.path.join(OUT, "launches_%s.csv" % tag), os.path.join(PROF, RND + "_launches.csv"))
main = summary(os.path.join(OUT, "prof_%s.ncu-rep" % tag))
open(os.path.join(PROF, RND + "_ncu_summary.txt"), "w").write(main)
if os.path.exists(os.path.join(OUT, "prof_other_%s.ncu-rep" % tag)):
    other = summary(os.path.join(OUT, "prof_other_%s.ncu-rep" % tag))
    if dbk:
        other += summary(os.path.join(OUT, "prof_%s.ncu-rep" % dbk))
    open(os.path.join(PROF, RND + "_other_kernels_summary.txt"), "w").write(other)

# DRAM traffic per picture of the capture (pictures per launch = bench default of the capture command)
pics = line["config"]["pics_per_step_per_gpu"]
blocks = re.split(r"^===== ", main, flags=re.M)[1:]
tr = {"source": "profiles/%s_ncu_summary.txt (ncu --set full, %d pictures per launch)" % (RND, pics)}
res_bytes, n_res = 0.0, 0
for b in blocks:
    name = b.splitlines()[0]
    rd = float(re.search(r"dram__bytes_read.sum\s+([\d.]+)", b).group(1)) * 1e6
    wr = float(re.search(r"dram__bytes_write.sum\s+([\d.]+)", b).group(1)) * 1e6
    if "residual_kernel" in name:
        res_bytes += rd + wr
        n_res += 1
    elif "sao_kernel" in name:
        tr["sao_kernel"] = {"dram_bytes_per_picture": (rd + wr) / pics, "algorithmic_bytes_per_picture": 49766400}
alg = line["roofline"]["alg_bytes_per_launch"] / pics if line["roofline"]["kernel"] == "residual_kernel" else None
tr["residual_kernel"] = {"dram_bytes_per_picture": res_bytes / pics, "algorithmic_bytes_per_picture": alg,
                         "instances": n_res}
json.dump(tr, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)

# SASS of the hot kernels
lib = os.path.join(REPO, "p265_b200", "libp265b200.so")
hist = []
for fn, out in (("_ZN4p26515residual_kernelILi0ELi2EEEvNS_10KernelArgsE", "sass_residual_kernel_32x32_sf_replicated.txt"),
                ("_ZN4p26515residual_kernelILi1ELi2EEEvNS_10KernelArgsE", "sass_residual_kernel_16x16_sf_replicated.txt"),
                ("_ZN4p26510sao_kernelItLb0ELi6EEEvNS_7SaoArgsE", "sass_sao_kernel_u16.txt"),
                ("_ZN4p26514deblock_kernelItLb1EEEvNS_7DbkArgsE", "sass_deblock_kernel_u16_packed.txt")):
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
    sass = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", sass)
    sass = "\n".join(l.rstrip() for l in sass.splitlines() if l.strip())
    open(os.path.join(PROF, out), "w").write(sass + "\n")
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", sass, flags=re.M)
    from collections import Counter
    c = Counter(ops)
    hist.append("== profiles/%s\n    " % out + "     ".join("%d %s" % (v, k) for k, v in c.most_common(14)))
open(os.path.join(PROF, "sass_opcode_histograms.txt"), "w").write("\n".join(hist) + "\n")
print("profiles/ updated from tag", tag, "value", line["value"], "residual frac", line["roofline"]["frac"])
