#!/usr/bin/env python
"""bench.py -- mel-frames/s and RTF of JETSGenerator.forward (PromptTTS AM + HiFi-GAN).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision fp32|tf32|bf16|fp32_ffma]
                    [--no-experiments]

Workload (BASELINE.json configs[1], the configuration `metric` is quoted on): batch = 1,
one 100-phoneme utterance (seed 1234 -> 537 mel frames = 8.6 s of 16 kHz audio), fp32,
full acoustic model + vocoder, seeded synthetic weights of the released architecture.
A "step" is one forward() over that batch.

* default arm: the engine (emotivoice_b200 -> libemotivoice_b200.so, sm_100a kernels).
  `value` is measured with inputs resident in HBM; `e2e` with inputs in pinned host memory
  (H2D inside the timed region) and the waveform read back to pinned host memory (D2H).
* --impl reference: the reference's algorithm on the host CPU (oracle/jets_oracle.py, the
  torch-CPU restatement pinned bit-exactly to the unmodified reference; the reference tree
  itself is Python and does not travel to the GPU box), all host threads.
At N=1, after the headline numbers are final, an "experiments" block re-measures the workload in the other precision modes
and the opt-in modes (EV_PDL, EV_AUTOTUNE, EV_FUSE_RES) plus the style encoder, each in a separate process under a timeout.
Under torchrun every rank runs the same per-GPU workload (weak scaling) after a one-time
NCCL weight broadcast from rank 0; time = max over ranks (CUDA events).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_PHONEMES = 100
SR, HOP = 16000, 256
WORKLOAD = "cfg2: batch=1, 100-phoneme utterance (seed 1234), fp32, PromptTTS AM + HiFi-GAN"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            sm, smax, reasons = [], [], set()
            with open(self.path) as f:
                for line in f:
                    p = [x.strip() for x in line.split(",")]
                    if len(p) < 9:
                        continue
                    try:
                        sm.append(float(p[1]))
                        smax.append(float(p[2]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            if sm:
                load = [x for x in sm if x > 0.5 * max(sm)] or sm
                out = dict(sm_mhz=statistics.median(load), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
            os.unlink(self.path)
        except Exception:
            pass
        return out


def usable_cpus():
    """CPUs this process may actually run on: affinity mask capped by the cgroup CPU quota
    (os.cpu_count() reports the whole host and oversubscribes a quota-limited container)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, per = f.read().split()
        if q != "max":
            n = min(n, max(1, int(math.ceil(float(q) / float(per)))))
    except Exception:
        pass
    return max(1, n)


def pick_cpu_threads(sd, conf):
    """The reference's own hint is `torch.set_num_threads(4)  # faster` (inference_tts.py:186):
    more threads is not monotonically better for these small convolutions.  Calibrate on a short
    vocoder-only sample and keep the fastest thread count <= the usable CPUs, so the CPU baseline
    is the best the host can do, not an oversubscribed one."""
    from emotivoice_b200 import synth
    from oracle import jets_oracle as O
    n = usable_cpus()
    cands = sorted({c for c in (4, 8, 16, 32, 64, n) if c <= n} or {n})
    mel = synth.make_mel(1, 48, seed=3)
    best, best_t, log = cands[0], float("inf"), {}
    for c in cands:
        torch.set_num_threads(c)
        O.vocoder(sd, conf.model, mel)
        t0 = time.perf_counter()
        O.vocoder(sd, conf.model, mel)
        dt = time.perf_counter() - t0
        log[c] = round(dt * 1e3, 1)
        if dt < best_t:
            best, best_t = c, dt
        if dt > 3 * best_t:
            break
    torch.set_num_threads(best)
    return best, n, log


def cpu_reference_run(steps, warmup):
    """The reference's algorithm on the host CPU (oracle port), best thread count <= usable CPUs."""
    from emotivoice_b200.config import default_config
    from emotivoice_b200 import synth
    from oracle import jets_oracle as O
    conf = default_config()
    sd = synth.make_state_dict(conf)
    cores, usable, calib = pick_cpu_threads(sd, conf)
    batch = synth.make_batch([N_PHONEMES], seed=synth.SEED)
    frames = 0
    for _ in range(max(1, warmup)):
        frames = int(O.jets_forward(sd, conf, **batch)["dec_outputs"].shape[1])
    t0 = time.perf_counter()
    for _ in range(steps):
        O.jets_forward(sd, conf, **batch)
    dt = (time.perf_counter() - t0) / steps
    return dict(frames=frames, sec_per_step=dt, fps=frames / dt, cores=cores, usable=usable, calib=calib)


def reference_arm(args, rank):
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(1, args.warmup)   # ~1 s of host CPU work per step
    r = cpu_reference_run(steps, warm)
    audio_s = r["frames"] * HOP / SR
    line = {
        "impl": "reference", "metric": "mel_frames_per_sec", "value": r["fps"], "unit": "mel-frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": r["sec_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames": r["frames"], "audio_seconds": audio_s, "device": "host CPU"},
        "rtf": r["sec_per_step"] / audio_s, "x_realtime": audio_s / r["sec_per_step"],
        "cpu_baseline": {"value": r["fps"], "unit": "mel-frames/s", "cores": r["cores"], "kind": "port",
                         "sample": "%d full forward passes of the bench workload (oracle/jets_oracle.py, torch %s CPU, %d threads "
                                   "= fastest of the calibration %s ms on a 48-frame vocoder sample; %d usable CPUs)"
                                   % (steps, torch.__version__, r["cores"], json.dumps(r["calib"]), r["usable"])},
        "e2e": {"value": r["fps"], "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def parity_vs_reference_fixture(out):
    """The bench workload is the committed fixture tests/golden/b1_t100.npz (inputs + outputs of the
    UNMODIFIED reference, oracle/make_golden.py): report the engine's distance from it."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "b1_t100.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    mel, wav, dur = torch.from_numpy(z["mel"]), torch.from_numpy(z["wav"]), torch.from_numpy(z["durations"])
    m, w = out["dec_outputs"].cpu(), out["wav_predictions"].cpu()
    ok = bool(torch.equal(out["log_duration_predictions"].cpu(), dur))
    res = {"reference": "tests/golden/b1_t100.npz (unmodified reference, CPU fp32)", "durations_identical": ok}
    if ok and m.shape == mel.shape:
        res["mel_max_abs_err_over_max_abs"] = float((m - mel).abs().max() / mel.abs().max())
        res["wav_rms_err_over_rms"] = float((w - wav).double().pow(2).mean().sqrt() / wav.double().pow(2).mean().sqrt())
    return res


def probe(args):
    """Child mode of `experiments()`: a short B=1 measurement of the same workload (3 warm-up + 10 timed forwards, L2
    flushed, CUDA events) in whatever mode the environment / --precision select; prints one small JSON object."""
    import __graft_entry__  # noqa: F401
    from emotivoice_b200.config import default_config
    from emotivoice_b200 import synth, _abi
    from emotivoice_b200.modules import JETSGenerator
    dev = torch.device("cuda", 0)
    conf = default_config()
    model = JETSGenerator(conf).to(dev)
    model.load_state_dict(synth.make_state_dict(conf))
    model.eval()
    model.precision = args.precision
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    if args.probe_workload == "voc":
        # cfg4 point (BASELINE.json configs[3]): vocoder only, batch 8 x 1024 frames; layer-granular traffic accounting of SURVEY s8d
        B, F = 8, 1024
        mel = synth.make_mel(B, F, seed=B * 7 + F).to(dev)
        for _ in range(2):
            model.generator(mel)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in ev:
            flush_buf.zero_()
            a.record()
            model.generator(mel)
            b.record()
        torch.cuda.synchronize()
        sec = sum(a.elapsed_time(b) for a, b in ev) / len(ev) * 1e-3
        peaks = _peaks()
        print(json.dumps({"workload": "vocoder only, B=8, F=1024", "ms_per_step": sec * 1e3, "mel_frames_per_sec": B * F / sec,
                          "tflops_algorithmic": B * F * 614105088.0 / sec / 1e12, "gbs_layer_granular": B * F * 5010752.0 / sec / 1e9,
                          "frac_hbm_layer_granular": B * F * 5010752.0 / sec / 1e9 / peaks["hbm_gbs"]}))
        return
    batch = {k: v.to(dev) for k, v in synth.make_batch([N_PHONEMES], seed=synth.SEED).items()}
    for _ in range(3):
        out = model(**batch)
    torch.cuda.synchronize()
    l0 = _abi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in ev:
        flush_buf.zero_()
        a.record()
        out = model(**batch)
        b.record()
    torch.cuda.synchronize()
    print(json.dumps({"ms_per_step": sum(a.elapsed_time(b) for a, b in ev) / len(ev),
                      "gpu_launches_per_step": (_abi.launch_count() - l0) / len(ev), "parity": parity_vs_reference_fixture(out)}))


def experiments(budget_s=330.0):
    """Opt-in modes measured AFTER the headline numbers are final, each in its own process under a timeout, so a failure
    or a hang in an experimental path cannot touch `value` / `e2e`.  Reported under "experiments"; never part of them."""
    runs = [("tf32", ["--precision", "tf32"], {}), ("bf16", ["--precision", "bf16"], {}),
            ("fp32+pdl", ["--precision", "fp32"], {"EV_PDL": "1"}),
            ("fp32+pdl_all", ["--precision", "fp32"], {"EV_PDL": "2"}),
            ("fp32+autotune", ["--precision", "fp32"], {"EV_AUTOTUNE": "2"}),
            ("fp32+fuse_res", ["--precision", "fp32"], {"EV_FUSE_RES": "1"}),
            ("fp32+all", ["--precision", "fp32"], {"EV_PDL": "2", "EV_AUTOTUNE": "1", "EV_FUSE_RES": "1"})]
    res = {"note": "opt-in / secondary modes of the same B=1 workload, 10 timed steps each, separate processes; not part of value or e2e"}
    t_end = time.time() + budget_s

    def child(cmd, env_extra, timeout):
        left = t_end - time.time()
        if left < 20:
            return {"skipped": "time budget of the experiments block spent"}
        try:
            r = subprocess.run(cmd, env=dict(os.environ, **env_extra), capture_output=True, text=True, timeout=min(timeout, left))
        except subprocess.TimeoutExpired:
            return {"error": "timeout"}
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout)[-300:])}
        d = json.loads(lines[-1])
        tune = [ln[len("[ev autotune] "):] for ln in r.stderr.splitlines() if ln.startswith("[ev autotune]")]
        if tune:
            d["autotune_log"] = tune
        return d

    for name, flags, env in runs:
        res[name] = child([sys.executable, os.path.abspath(__file__), "--probe"] + flags, env, 90)
    for name, prec, env in (("voc_b8_f1024_fp32", "fp32", {}), ("voc_b8_f1024_fp32+fuse_res", "fp32", {"EV_FUSE_RES": "1"}),
                            ("voc_b8_f1024_tf32", "tf32", {}), ("voc_b8_f1024_tf32+fuse_res", "tf32", {"EV_FUSE_RES": "1"})):
        res[name] = child([sys.executable, os.path.abspath(__file__), "--probe", "--probe-workload", "voc", "--precision", prec], env, 90)
    res["style_encoder"] = child([sys.executable, os.path.join(ROOT, "tools", "style_bench.py"), "--steps", "20"], {}, 150)
    return res


def dominant_kernel_roofline(lib, dev, frames, peaks, flush, precision="fp32"):
    """conv1d_tm_kernel on the single most expensive layer shape of the step: the k=11
    ResBlock convolutions of HiFi-GAN stage 2 (C=128, L=64*F; 22% of all FLOPs).  Timed live
    with CUDA events on the launching stream, L2 flushed before every launch."""
    from emotivoice_b200 import _abi
    C, K, L = 128, 11, 64 * frames
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, L, C, generator=g).to(dev)
    w = torch.randn(K, C, C, generator=g) / math.sqrt(C * K)
    use_tc = precision in ("fp32", "tf32", "bf16")
    if use_tc:
        from emotivoice_b200 import packing
        w = packing.to_tc16_layout(w) if precision == "bf16" else packing.to_tc_layout(w)
    w = w.to(dev)
    split3 = {"fp32": 1, "tf32": 0, "bf16": 2}.get(precision, 0)
    b = torch.randn(C, generator=g).to(dev)
    res = torch.randn(1, L, C, generator=g).to(dev)
    out = torch.empty(1, L, C, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    times = []
    for i in range(13):
        flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if use_tc:
            _abi.check(lib.ev_op_conv1d_tc(x.data_ptr(), w.data_ptr(), split3, b.data_ptr(), 0, res.data_ptr(), out.data_ptr(), 1, L, C, C,
                                           K, 1, None, 1, _abi.ACT_LRELU, 0.1, _abi.ACT_NONE, _abi.ACC_STORE, 1.0, None, 0, st))
        else:
            _abi.check(lib.ev_op_conv1d(x.data_ptr(), w.data_ptr(), b.data_ptr(), 0, res.data_ptr(), out.data_ptr(), 1, L, C, C,
                                        K, 1, None, 1, _abi.ACT_LRELU, 0.1, _abi.ACT_NONE, _abi.ACC_STORE, 1.0, st))
        e1.record()
        e1.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1) * 1e-3)
    t = statistics.mean(times)
    flops = 2.0 * L * C * C * K
    alg_bytes = 4.0 * (L * C * 3) + 4.0 * K * C * C
    achieved = flops / t / 1e12
    ffma_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    if use_tc:
        mma_mult = 3 if split3 == 1 else 1
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")   # dram bytes/launch from the committed ncu --set full capture
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(precision, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        return {
            "kernel": "conv1d_tc_kernel<%s> (tcgen05 kind::tf32, %s; HiFi-GAN stage-2 ResBlock conv, C=128, k=11, L=%d)"
                      % (split3, ("1xTF32", "3xTF32 fp32 emulation", "bf16 operands")[split3], L),
            "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
            "peak_source": "%s bf16 burst (MEASURED_PEAKS.json). `achieved` counts ALGORITHMIC flops (2*L*Cin*Cout*k); the tensor "
                           "pipe executes %dx that in tf32 MMAs at half the bf16 rate, so tensor-pipe occupancy ~ %d*frac"
                           % (peaks["source"], mma_mult, 2 * mma_mult),
            "tensor_pipe_frac_est": (1 if split3 == 2 else 2 * mma_mult) * achieved / peaks["bf16_tflops"],
            "flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes,
            "avg_launch_ms": t * 1e3, "hbm_gbs_at_algorithmic_bytes": alg_bytes / t / 1e9,
            "hbm_frac_at_algorithmic_bytes": alg_bytes / t / 1e9 / peaks["hbm_gbs"],
        }
    return {
        "kernel": "conv1d_tm_kernel<16,2,8> (HiFi-GAN stage-2 ResBlock conv, C=128, k=11, L=%d)" % L,
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_tflops"], "traffic": None,
        "peak_source": "%s bf16 burst (MEASURED_PEAKS.json)" % peaks["source"],
        "note": "round-1 kernel is an fp32 FFMA implicit GEMM (exact fp32 parity path); against the fp32 FFMA "
                "ceiling of 148 SMs x 128 lanes x 2 x 1.965 GHz = %.1f TFLOP/s it reaches frac_fp32_ffma" % ffma_peak,
        "frac_fp32_ffma": achieved / ffma_peak,
        "flops_per_launch": flops, "algorithmic_bytes_per_launch": alg_bytes,
        "avg_launch_ms": t * 1e3, "hbm_gbs_at_algorithmic_bytes": alg_bytes / t / 1e9,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default=os.environ.get("EV_PRECISION", "fp32"), choices=["fp32", "tf32", "bf16", "fp32_ffma"])
    ap.add_argument("--probe", action="store_true", help="internal: short child measurement for the experiments block")
    ap.add_argument("--probe-workload", default="b1", choices=["b1", "voc"])
    ap.add_argument("--no-experiments", action="store_true", help="skip the opt-in-mode block measured after the headline")
    args = ap.parse_args()
    if args.probe:
        probe(args)
        return
    if args.impl == "engine":
        args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import __graft_entry__  # noqa: F401  (sys.path)
    from emotivoice_b200 import build as _build
    _build.build(verbose=False)          # no-op when the in-tree .so is current; file-locked, so every rank may call it
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    from emotivoice_b200.config import default_config
    from emotivoice_b200 import synth, _abi, runner
    from emotivoice_b200.modules import JETSGenerator

    conf = default_config()
    sd = synth.make_state_dict(conf) if rank == 0 else None
    bcast_ms = None
    if world > 1:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sd = runner.broadcast_state_dict(sd, conf, dev, src=0)     # one-time NCCL weight broadcast
        torch.cuda.synchronize()
        bcast_ms = (time.perf_counter() - t0) * 1e3
    model = JETSGenerator(conf).to(dev)
    model.load_state_dict(sd)
    model.eval()
    model.precision = args.precision
    lib = _abi.load()

    batch_cpu = synth.make_batch([N_PHONEMES], seed=synth.SEED)
    batch_dev = {k: v.to(dev) for k, v in batch_cpu.items()}
    batch_pin = {k: v.pin_memory() for k, v in batch_cpu.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in batch_pin.values())

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2

    def flush():
        flush_buf.zero_()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -----------------------------------------------------------------------
    out = None
    for _ in range(args.warmup):
        out = model(**batch_dev)
    torch.cuda.synchronize()
    parity = parity_vs_reference_fixture(out) if rank == 0 else None
    frames = int(out["dec_outputs"].shape[1])
    n_samples = int(out["wav_predictions"].shape[-1])
    wav_pin = torch.empty((1, 1, n_samples), dtype=torch.float32).pin_memory()
    audio_s = frames * HOP / SR

    # ---- timed: inputs resident in HBM -------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = _abi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush()
        ev[s][0].record()
        out = model(**batch_dev)
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = _abi.launch_count() - l0
    dev_s = sum(a.elapsed_time(b) for a, b in ev) * 1e-3

    # ---- timed: end to end (pinned host inputs -> device, waveform -> pinned host) -------------------
    for _ in range(2):
        o = model(**{k: v.to(dev, non_blocking=True) for k, v in batch_pin.items()})
        wav_pin.copy_(o["wav_predictions"], non_blocking=True)
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush()
        ev2[s][0].record()
        o = model(**{k: v.to(dev, non_blocking=True) for k, v in batch_pin.items()})
        wav_pin.copy_(o["wav_predictions"], non_blocking=True)
        ev2[s][1].record()
    barrier()
    e2e_s = sum(a.elapsed_time(b) for a, b in ev2) * 1e-3
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks ------------------------------------------------------------------
    t = torch.tensor([dev_s, e2e_s, float(frames)], dtype=torch.float64, device=dev)
    if dist is not None:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_s, e2e_s, total_frames = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        total_frames = float(frames)

    if rank == 0:
        peaks = _peaks()
        value = total_frames * args.steps / dev_s
        e2e_val = total_frames * args.steps / e2e_s
        ms_step = dev_s / args.steps * 1e3
        roof = dominant_kernel_roofline(lib, dev, frames, peaks, flush, args.precision)
        line = {
            "metric": "mel_frames_per_sec", "value": value, "unit": "mel-frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": {"fp32": "fp32 (fp32 storage; GEMM-shaped layers on tcgen05 as 3xTF32 fp32 emulation, fp32 accumulation in TMEM; "
                              "everything else fp32 FFMA)",
                      "tf32": "tf32 (fp32 storage; decoder+vocoder GEMMs on tcgen05 with tf32 operands rounded to nearest, fp32 "
                              "accumulation; duration prefix 3xTF32)",
                      "bf16": "bf16 (fp32 storage; decoder+vocoder GEMMs on tcgen05 kind::f16 with bf16 operands, fp32 accumulation; "
                              "duration prefix 3xTF32)",
                      "fp32_ffma": "fp32 (FFMA kernels, no tensor cores)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "precision": args.precision, "frames_per_step_per_gpu": frames, "audio_seconds_per_step_per_gpu": audio_s,
                       "l2": "flushed between timed steps (256 MiB write, outside the event pairs)",
                       "timing": "sum of per-step CUDA-event pairs on the launching stream, max over ranks",
                       "weights": "seeded synthetic, 53.3 M params fp32",
                       "weight_broadcast_ms": bcast_ms},
            "rtf": (dev_s / args.steps) / audio_s, "x_realtime": audio_s / (dev_s / args.steps),
            "wall_ms_per_step_incl_flush": wall / args.steps * 1e3,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "mel-frames/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": n_samples * 4, "ms_per_step": e2e_s / args.steps * 1e3,
                    "x_realtime": audio_s / (e2e_s / args.steps)},
            "gpu_launches": int(launches),
            "parity": parity,
            "roofline": roof,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(steps=8, warmup=1)
            line["cpu_baseline"] = {"value": r["fps"], "unit": "mel-frames/s", "cores": r["cores"], "kind": "port",
                                    "sample": "8 full forward passes of the same workload on the host CPU (oracle/jets_oracle.py, "
                                              "torch CPU, %d threads = fastest of calibration %s ms; %d usable CPUs); %.0f ms each"
                                              % (r["cores"], json.dumps(r["calib"]), r["usable"], r["sec_per_step"] * 1e3)}
        if world == 1 and not args.no_experiments and os.environ.get("EV_BENCH_EXPERIMENTS", "1") != "0":
            try:
                line["experiments"] = experiments()
            except Exception as e:      # never let the secondary block cost the headline line
                line["experiments"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
