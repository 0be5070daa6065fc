"""Fill the {{...}} measurement placeholders of DESIGN.md from the committed bench lines under profiles/ (run after the final measurement call)."""
import json
import os

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ld = lambda f: json.loads([l for l in open(os.path.join(R, "profiles", f)) if l.startswith("{")][-1])
j = ld("r2_bench_netflix_1gpu.json")
j2, j8 = ld("r2_bench_synthetic_2gpu.json"), ld("r2_bench_synthetic_8gpu.json")
c = j["configs"]
f = j["roofline"]["families_ms"]; fr = j["roofline"]["families_frac"]; fg = j["roofline"]["families_gbs"]
k = lambda v: f"{v / 1e6:.3f} M" if v >= 1e6 else f"{v / 1e3:.1f} k"
syn1 = c["synthetic_1gpu"]
rep = {
    "NF_MS": f"{j['ms_per_step']:.3f}", "NF_V": k(j["value"]), "NF_E2E": k(j["e2e"]["value"]), "NF_EVAL": k(j["eval"]["value"]),
    "NFH_MS": f"{c['netflix_hoisted']['ms_per_step']:.3f}", "NFH_V": k(c["netflix_hoisted"]["value"]), "NFH_E2E": k(c["netflix_hoisted"]["e2e"]["value"]),
    "NFD_MS": f"{c['netflix_hoisted_device_sampler']['ms_per_step']:.3f}", "NFD_V": k(c["netflix_hoisted_device_sampler"]["value"]),
    "NFD_E2E": k(c["netflix_hoisted_device_sampler"]["e2e"]["value"]),
    "ML_MS": f"{c['movielens']['ms_per_step']:.3f}", "ML_V": k(c["movielens"]["value"]), "ML_E2E": k(c["movielens"]["e2e"]["value"]), "ML_EVAL": k(c["movielens"]["eval"]["value"]),
    "SYN1_MS": f"{syn1['ms_per_step']:.1f}", "SYN2_MS": f"{j2['ms_per_step']:.1f}", "SYN8_MS": f"{j8['ms_per_step']:.1f}",
    "SYN1_V": k(syn1["value"]), "SYN2_V": k(j2["value"]), "SYN8_V": k(j8["value"]),
    "SYN2_X": f"{j2['speedup_vs_1gpu']:.2f}", "SYN8_X": f"{j8['speedup_vs_1gpu']:.2f}",
    "SYN1_EVAL": k(syn1["eval"]["value"]), "SYN8_EVAL": k(j8["eval"]["value"]), "SYN1_EVAL_MS": f"{syn1['eval']['ms']:.1f}", "SYN1_TF": f"{syn1['eval']['tensor_tflops_useful']:.0f}",
    "CPU_MS": f"{j['cpu_baseline']['ms_per_step']:.0f}", "CPU_V": k(j["cpu_baseline"]["value"]), "CPU_EVAL": f"{j['cpu_baseline']['eval']['value']:.0f}",
    "GT_MS": f"{j['gpu_torch_baseline']['ms_per_step']:.1f}", "GT_V": k(j["gpu_torch_baseline"]["value"]), "GT_EVAL": f"{j['gpu_torch_baseline']['eval']['value']:.0f}",
    "X_CPU": f"{j['e2e']['value'] / j['cpu_baseline']['value']:.0f}", "X_GT": f"{j['gpu_torch_baseline']['ms_per_step'] / j['ms_per_step']:.0f}",
    "F_WG": f"{f['proj_wgrad']:.3f}", "G_WG": f"{fg['proj_wgrad']:.0f}", "R_WG": f"{fr['proj_wgrad']:.2f}", "F_FW": f"{f['proj_fwd']:.3f}", "R_FW": f"{fr['proj_fwd']:.2f}",
    "F_SB": f"{f['spmm_bwd']:.3f}", "R_SB": f"{fr['spmm_bwd']:.2f}", "F_SF": f"{f['spmm_fwd']:.3f}", "R_SF": f"{fr['spmm_fwd']:.2f}",
    "F_FB": f"{f['fuse_bwd']:.3f}", "F_LH": f"{f['loss_heads']:.3f}", "F_FF": f"{f['fuse_fwd']:.3f}", "F_AD": f"{f['adamw']:.3f}", "R_AD": f"{fr['adamw']:.2f}",
}
p = os.path.join(R, "DESIGN.md")
s = open(p).read()
for a, b in rep.items():
    s = s.replace("{{" + a + "}}", b)
left = [w for w in s.split("{{")[1:]]
open(p, "w").write(s)
print("filled", len(rep), "placeholders; unfilled:", [w.split("}}")[0] for w in left])
