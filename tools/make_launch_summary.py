"""profiles/r02_launch_summary.md from the committed evidence files (bench line, launch lists, kernel times)."""
import json, os, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/"
run = lambda f: subprocess.run(["python", "tools/launch_summary.py", f], capture_output=True, text=True, cwd=root).stdout.split("\n")
tr = run("profiles/r02_launches_train_steps3_objects64.csv")
inf = run("profiles/r02_launches_infer_steps3_objects64.csv")
d = json.load(open(root + "profiles/r02_bench_n1.json"))
lt = open(root + "profiles/r02_layer_times.txt").read()
md = ["# Round 2 launch summary (B200, final build)\n"]
md.append("Sources: `profiles/r02_launches_train_steps3_objects64.csv` / `r02_launches_infer_steps3_objects64.csv` = `ncu --metrics gpu__time_duration.sum --clock-control none --csv` of `tools/step_once.py 64 3 [infer]` (3 eager steps of the bench's configs[1] batch, 1 208 829 cells, incl. the one-off graph layout launches; per-launch times are cold-cache and serialised, so the SHARES are what agrees with the live bench, not the absolutes); `profiles/r02_ncu_layer_kernels.csv` = `ncu --set full` of `tools/exp_layer_one.py` per layer shape (DRAM bytes, unit utilisation per kernel); `profiles/r02_layer_times.txt` = the same kernels timed with CUDA events (no profiler); `profiles/r02_bench_n1.json` = `python bench.py`; `profiles/r02_bench_n{2,4,8}.json` = the partitioned runs; `profiles/r02_sass_histogram.txt` = SASS opcodes per kernel.\n")
md.append("## Live bench (CUDA events, no profiler)\n")
md.append("* step %.2f ms = %.3e cells/s resident (CUDA graph replay; eager %.2f ms), e2e %.3e cells/s (%.2f ms per step, %d MB uploaded per step), step_hbm_frac %.4f" % (d["ms_per_step"], d["value"], d["ms_per_step_eager"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"] // 10**6, d["step_hbm_frac"]))
r = d["roofline"]
md.append("* dominant C-ABI call `%s`: %.4f ms per launch, %.1f GB/s of %.1f = %.4f; DRAM traffic %.3f GB per launch against %.3f GB algorithmic (%.2fx; `%s`); share of the step %.3f" % (r["kernel"], r["kernel_ms_per_launch"], r["achieved"], r["peak"], r["frac"], r["traffic"] / 1e9, r["algorithmic_bytes_per_launch"] / 1e9, r["traffic"] / r["algorithmic_bytes_per_launch"], r["traffic_source"], r["kernel_share_of_step"]))
md.append("* inference %.3f ms per 1.21 M-cell pass (hbm_frac %.4f); cfg1 %.3f ms (%.3e cells/s, parity %.1e); cfg3 %.2f ms for 6.75 M cells; 67.1 M-cell scene %.1f ms; wide [128,256,512,1024] training %.1f ms per 302 602 cells = %.1f algorithmic TFLOP/s; Updated-filter training %.1f ms per 2.1 M cells" % (d["inference"]["ms"], d["inference"]["hbm_frac"], d["cfg1_inference"]["ms"], d["cfg1_inference"]["cells_per_s"], d["cfg1_inference"]["parity_max_rel"], d["cfg3_inference"]["ms"], d["scene_inference"]["ms"], d["wide_training"]["ms_per_step"], d["wide_training"]["algorithmic_tflops"], d["updated_training"]["ms_per_step"]))
md.append("* parity of the bench's own first step against the oracle: logits max rel %.2e (tolerance 1e-4), loss %.8f vs %.8f, label flips off ties %d" % (d["parity"]["logits_max_rel"], d["parity"]["loss_cuda"], d["parity"]["loss_oracle"], d["parity"]["label_flips_off_ties"]))
md.append("* clocks: %s\n" % json.dumps(d["clocks"]))
md.append("Per C-ABI call, ms per step (live, CUDA events around every call):\n\n| call | ms / step | launches / step |\n|---|---|---|")
md += ["| `%s` | %.4f | %d |" % (k[0], k[1], k[2]) for k in d["kernels"]]
md.append("\n## Kernel times per layer shape (CUDA events, 604 913 cells)\n\n```\n" + lt + "```\n")
md.append("## Training step launch list under ncu (3 steps)\n\n```\n" + "\n".join(tr[:34]) + "\n```\n")
md.append("## Inference launch list under ncu (3 passes)\n\n```\n" + "\n".join(inf[:14]) + "\n```\n")
tc = ("dw2_tc", "dense2_tc", "gather_tc", "dwe_tc", "layer_tc", "dw_tc_kernel")
md.append("Share check: the ncu list gives the tensor-core layer kernels %.1f %% of the step's kernel time; the live per-call table gives its top 12 (the same kernels) %.1f %%." % (
    100 * sum(float(l.split()[-3]) for l in tr[1:] if any(t in l for t in tc)) / float(tr[0].split()[-2]),
    100 * sum(k[1] for k in d["kernels"]) / d["kernel_ms_per_step"]))
open(root + "profiles/r02_launch_summary.md", "w").write("\n".join(md) + "\n")
print("wrote profiles/r02_launch_summary.md")
