"""Builds profiles/r02_summary.md from the committed round-2 measurement files (so every number in it can be traced to a file).
   python tools/make_r02_summary.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
HBM = 6541.8


def jl(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return []
    return [json.loads(l) for l in open(path) if l.startswith("{")]


def last_json(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    out = ["# Round 2 measurements (one B200, sm_100a, 1965 MHz, no throttle reasons; every table cites its file)", ""]
    b = last_json("r02_bench_n1.json")
    r1 = json.load(open(os.path.join(P, "r01_bench_final_fp32.json")))
    if b:
        out += ["## Headline: `bench.py --gpus 1 --steps 20 --warmup 5` (`r02_bench_n1.json`; round 1: `r01_bench_final_fp32.json`)", "",
                "cfg2: batch 1, one 100-phoneme utterance per step (distinct utterances), fp32 mode.", "",
                "| | round 1 | round 2 |", "|---|---|---|",
                "| device time per utterance (`ms_per_step`) | %.2f ms | **%.2f ms** |" % (r1["ms_per_step"], b["ms_per_step"]),
                "| `value` (mel-frames/s, inputs resident) | %.0f | **%.0f** |" % (r1["value"], b["value"]),
                "| `e2e` (host data -> pinned -> H2D -> forward -> wav D2H) | %.2f ms | **%.2f ms** = %.0f frames/s |" % (r1["e2e"]["ms_per_step"], b["e2e"]["ms_per_step"], b["e2e"]["value"]),
                "| x real time (device) | %.0f | **%.0f** |" % (r1["x_realtime"], b["x_realtime"]),
                "| kernel launches per utterance | 202 | %d |" % round(b["gpu_launches"] / b["steps"] / b["n_gpus"]),
                "| parity vs the unmodified reference's fixture | mel 5.4e-6, wav 7.8e-7 | mel %.1e, wav %.1e (durations identical) |" % (b["parity"]["mel_relmax"], b["parity"]["wav_relrms"]),
                "| CPU baseline in the same run (oracle port, %d threads) | 1.34-1.36 k frames/s | %.0f frames/s |" % (b["cpu_baseline"]["cores"], b["cpu_baseline"]["value"]),
                ""]
        rf = b["roofline"]
        tr = json.load(open(os.path.join(P, "dominant_kernel_traffic.json"))).get("fp32", {})
        out += ["Dominant launch, timed live inside bench.py: %s: %.1f us, %.0f algorithmic TFLOP/s = **%.3f** of the measured bf16 peak (%.0f), "
                "executing 3x that in bf16 MMAs (tensor-pipe estimate %.2f; ncu of the same launch: tensor pipe %.1f %% active, %.1f us, DRAM traffic %.1f MB "
                "vs %.1f MB algorithmic -- the in-place outputs stay in L2; `%s`).  Round 1: one convolution of this stage at 0.083 (88 us); the same single "
                "convolution this round: 0.153 (41-46 us)." % (rf["kernel"], rf["ms"] * 1e3, rf["achieved"], rf["frac"], rf["peak"], rf.get("tensor_pipe_frac_est", 0),
                                                            tr.get("tensor_pipe_pct", 0), tr.get("duration_us", 0), (rf.get("traffic") or 0) / 1e6,
                                                            tr.get("algorithmic_bytes_per_launch", 0) / 1e6, tr.get("source", "").split(" ")[0]), ""]
        if "b1" in b:
            out += ["Fixture utterance (537 frames, 8.59 s of audio): %.2f ms = %.0fx real time (round 1: 6.28 ms)." % (b["b1"]["ms"], b["b1"]["x_rt"]), ""]
        if "b32" in b and "bf16" in b["b32"]:
            q = b["b32"]
            out += ["cfg3 (32 mixed EN/ZH utterances, %d valid of %d padded frames): bf16 %.1f ms = %.0f valid frames/s (parity of 3 items vs the oracle's "
                    "B=1 fp32 runs: durations identical, wav rel-rms <= %.1e, tolerance 2e-2); fp32 mode %.1f ms = %.0f frames/s.  Round 1: fp32 135 ms, tf32 86 ms."
                    % (q["valid_frames"], q["padded_frames"], q["bf16"]["ms"], q["bf16"]["valid_fps"], q["parity_bf16_vs_oracle_b1"]["wav_relrms_max"],
                       q.get("fp32", {}).get("ms", float("nan")), q.get("fp32", {}).get("valid_fps", float("nan"))), ""]
        if "voc" in b:
            out += ["cfg4 points inside the bench line (vocoder only; `hbm_frac` = 5,010,752 B/frame x frames / time / %.1f GB/s):" % HBM, "",
                    "| point | ms | mel-frames/s | algorithmic TFLOP/s | HBM fraction (layer-granular) |", "|---|---|---|---|---|"]
            for k, v in b["voc"].items():
                if isinstance(v, dict):
                    out.append("| %s | %.2f | %.0f | %.0f | %.3f |" % (k, v["ms"], v["fps"], v["tflops"], v["hbm_frac"]))
            out.append("")
        if "cfg5" in b:
            c = b["cfg5"]
            out += ["cfg5 block (256 utterances per GPU, B=32 buckets, pinned int16 D2H, wall clock): %.0f utterances/s, %.0f frames/s at N=1 "
                    "(host collate %.1f ms, trim + sha1 %.0f ms of %.0f ms)." % (c["utt_per_s"], c["fps"], c["host_collate_s"] * 1e3, c["host_finish_s"] * 1e3, c["wall_s"] * 1e3), ""]
        if "cfg5_strong" in b:
            c = b["cfg5_strong"]
            out += ["cfg5_strong (the same pipeline on a FIXED 2048-utterance corpus; strong scaling): %.0f utterances/s, %.0f frames/s at N=1 (%.2f s)."
                    % (c["utt_per_s"], c["fps"], c["wall_s"]), ""]
    for n in (2, 4, 8):
        bn = last_json("r02_bench_n%d.json" % n)
        if bn and b:
            out += ["N=%d (`r02_bench_n%d.json`): value %.0f frames/s (%.2fx N=1), e2e %.0f (%.2fx); cfg5 %.0f utt/s (%.2fx), digest of utterances 0..255 %s "
                    "(N=1: %s); cfg5_strong (2048 utterances, fixed) %.0f utt/s (%.2fx), digest over all of them %s (N=1: %s); weight distribution: raw %.0f ms + pack-and-blob %.0f ms (%.0f MB)."
                    % (n, n, bn["value"], bn["value"] / b["value"], bn["e2e"]["value"], bn["e2e"]["value"] / b["e2e"]["value"], bn["cfg5"]["utt_per_s"],
                       bn["cfg5"]["utt_per_s"] / b["cfg5"]["utt_per_s"], bn["cfg5"].get("digest_first256"), b["cfg5"].get("digest_first256"),
                       bn["cfg5_strong"]["utt_per_s"], bn["cfg5_strong"]["utt_per_s"] / b["cfg5_strong"]["utt_per_s"], bn["cfg5_strong"].get("digest"),
                       b["cfg5_strong"].get("digest"), bn["weights"]["raw_ms"], bn["weights"]["pack_and_blob_ms"], bn["weights"]["blob_mb"]), ""]

    sw = os.path.join(P, "r02_sweep_cfg4_with_corner.json")
    if os.path.exists(sw):
        d = json.load(open(sw))
        out += ["## cfg4 sweep incl. the B=128, F=4096 corner (`r02_sweep_cfg4_with_corner.json`, `tools/sweep.py --quick --corner`)", "",
                "| precision | B | F | ms | mel-frames/s | HBM fraction |", "|---|---|---|---|---|---|"]
        for r in d["cfg4"]:
            if "seconds" in r:
                out.append("| %s | %d | %d | %.2f | %.0f | %.3f |" % (r["precision"], r["batch"], r["frames"], r["seconds"] * 1e3, r["mel_frames_per_sec"], r["frac_hbm"]))
        out += ["", "(round 1 saturated at 230 k (fp32) / 403 k (tf32) frames/s = 0.17 / 0.31; `hbm_frac` uses SURVEY.md s8d's fp32 byte count for every "
                    "mode, so the bf16 rows -- whose activations are stored as bf16 -- move half those bytes: their DRAM utilisation is about half the figure.)", ""]

    for name, title in (("r02_layer_sweep_b8_f1024.jsonl", "B=8, F=1024 (the cfg4 point)"), ("r02_layer_sweep_b1_f537.jsonl", "B=1, F=537 (the cfg2 utterance)")):
        rows = jl(name)
        if not rows:
            continue
        out += ["## Per-layer microbenchmarks, %s (`%s`, `tools/layer_sweep.sh`): granule-planar kernel (gp:*) vs the round-1 time-major kernel" % (title, name), "",
                "L2 flushed before every launch; `GB/s` = layer-granular bytes (in + out + residual, weights once) / time; HBM peak %.0f GB/s." % HBM, "",
                "| mode | C_in | C_out | k | dil | us | TFLOP/s | GB/s | of HBM peak |", "|---|---|---|---|---|---|---|---|---|"]
        for r in rows:
            out.append("| %s | %d | %d%s | %d | %d | %.1f | %.0f | %.0f | %.2f |" % (r["mode"], r["Cin"], r["Cout"], (" (x%d phases)" % r["rate"]) if r["rate"] > 1 else "",
                                                                             r["K"], r["dil"], r["best_us"], r["tflops"], r["layer_gbs"], r["layer_gbs"] / HBM))
        out.append("")
    rows = jl("r02_fused_vs_unfused.jsonl")
    if rows:
        out += ["## Fused ResBlock layer vs the two launches it replaces (`r02_fused_vs_unfused.jsonl`, `tools/profile_pair.py`)", "",
                "| mode | C | k | dil | B | fused us | unfused us | speed-up | fused TFLOP/s |", "|---|---|---|---|---|---|---|---|---|"]
        for r in rows:
            out.append("| %s | %d | %d | %d | %d | %.1f | %.1f | **%.2f** | %.0f |" % (r["mode"], r["C"], r["K"], r["dil"], r["B"], r["fused_us"], r["unfused_us"], r["speedup"], r["fused_tflops"]))
        out += ["", "Round 1's fused kernel was 10-21 % SLOWER than its unfused path.  The engine fuses the 32-channel layers always and the 64-channel ones except "
                    "k = 11 at large batch (both paths are bitwise equal, so the choice may depend on the batch).  (Measured mid-round, before PDL and the grouped launches.)", ""]
    rows = jl("r02_attention_tc_vs_ffma.jsonl")
    if rows:
        out += ["## Attention: tcgen05 kernel vs the fp32 FFMA flash kernel (`r02_attention_tc_vs_ffma.jsonl`, `tools/profile_attn.py`)", "",
                "| B | L | mode | tcgen05 us | FFMA us | speed-up |", "|---|---|---|---|---|---|"]
        for r in rows:
            t, f = min(r["us"]["tc"][1:]), min(r["us"]["ffma"][1:])
            out.append("| %d | %d | %s | %.1f | %.1f | %.2f |" % (r["B"], r["L"], "3xTF32" if r["tc_mode"] else "tf32", t, f, f / t))
        out.append("")
    pk = os.path.join(P, "r02_packed_batch_layer_shapes.json")
    if os.path.exists(pk):
        d = json.load(open(pk))
        out += ["## Acoustic-model GEMM shapes at packed-batch size, time-major kernel conv1d_tc (`r02_packed_batch_layer_shapes.json`)", "", d["what"] + ".  " + d["peak_note"] + ".", "",
                "| mode | layer | us | algorithmic TFLOP/s | tensor-pipe fraction (executed) |", "|---|---|---|---|---|"]
        for r in d["rows"]:
            out.append("| %s | %s | %.1f | %.0f | %.2f |" % (r["mode"], r["layer"], r["us"], r["tflops_algorithmic"], r["tensor_pipe_frac_est"]))
        out += ["", "Round 1 (`r01_packed_batch_layer_shapes.json`): 3xTF32 0.34 / 0.62 / 0.52, tf32 0.13 / 0.36 / 0.30.  The >= 60 % target is met by the fp32-class "
                    "modes' conv-FFN layers (0.73 / 0.67 3xTF32, 0.63 / 0.56 bf16x3) and missed by the 1x modes (tf32 0.40, bf16 0.28): this kernel still stages "
                    "A through registers (the vocoder's granule-planar operand path was not ported to the acoustic model).", ""]
    for name, title in (("r02_pdl_modes.log", "Programmatic dependent launch (EV_PDL = 0 / 1 / 2; before the grouped launches)"),
                        ("r02_grouped_launches_b1.log", "Grouped launches on / off, fixture utterance, final code (`EV_VOC_GROUP`)"),
                        ("r02_dominant_launch.log", "The dominant launch standalone (`tools/profile_dominant.py`: the grouped stage-2 convolutions, L2 flushed)")):
        path = os.path.join(P, name)
        if os.path.exists(path):
            out += ["## %s (`%s`)" % (title, name), "", "```"] + [l.rstrip()[:230] for l in open(path).read().splitlines()] + ["```", ""]
    wc = jl("r02_warm_vs_cold_b1.jsonl")
    if wc:
        out += ["## Batch-1 layers: L2-cold vs weights-warm vs everything-warm (`r02_warm_vs_cold_b1.jsonl`): weight fetch is not the limiter", "",
                "| mode | C_in | C_out | k | L | cold us | weights warm us | all warm us |", "|---|---|---|---|---|---|---|---|"]
        key = lambda r: (r["mode"], r["Cin"], r["Cout"], r["K"], r["L"])
        tab = {}
        for r in wc:
            tab.setdefault(key(r), {})[r["warm"]] = r["best_us"]
        for k, v in tab.items():
            out.append("| %s | %d | %d | %d | %d | %.1f | %.1f | %.1f |" % (k + (v.get("none", 0), v.get("w", 0), v.get("all", 0))))
        out.append("")
    for name, per, title in (("r02_launches_b1_fp32.csv", 3, "one B=1 fp32 step (cfg2 fixture utterance)"), ("r02_launches_b32_bf16.csv", 2, "one B=32 bf16 step (cfg3 batch)")):
        path = os.path.join(P, name)
        if os.path.exists(path):
            txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), path, "--steps-by", "validate_inputs"], capture_output=True, text=True).stdout
            out += ["## ncu launch list of %s (`%s`; cold-cache, serialised: compare shares)" % (title, name), "", "```"] + txt.splitlines()[:16] + ["```", ""]
    open(os.path.join(P, "r02_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
