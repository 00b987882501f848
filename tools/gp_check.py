"""First-contact check of the granule-planar convolution kernel (csrc/conv1d_gp.cu) against the round-1 time-major
tensor-core kernel (bitwise in the fp32 / tf32 modes: same reduction order, same rounding) and a torch reference (bf16).
Run it under `timeout` before the test suite: a deadlock in a new mbarrier pipeline must not take pytest with it.

    timeout 180 python tools/gp_check.py [--quick]
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from emotivoice_b200 import _abi, build, layout, packing

CASES = [
    # B, L, Cin, Cout, K, dil, rate, use_res, acc, ragged
    (1, 128, 32, 32, 1, 1, 1, 0, 0, 0),
    (1, 128, 32, 32, 3, 1, 1, 0, 0, 0),
    (1, 300, 32, 32, 11, 5, 1, 1, 0, 0),
    (2, 517, 64, 64, 7, 3, 1, 1, 1, 1),
    (1, 70, 80, 512, 7, 1, 1, 0, 0, 0),          # conv_pre: C_in tail block
    (3, 700, 128, 128, 3, 1, 1, 1, 2, 1),
    (2, 900, 256, 256, 11, 1, 1, 1, 0, 1),
    (1, 200, 512, 2048, 3, 1, 8, 0, 0, 0),       # ups[0] polyphase
    (2, 1500, 128, 128, 3, 1, 2, 0, 0, 1),       # ups[2]: rate 2, 64 channels per phase
    (2, 3000, 64, 64, 3, 1, 2, 0, 0, 1),         # ups[3]: rate 2, 32 channels per phase
    (3, 70000, 32, 32, 11, 5, 1, 1, 2, 1),       # persistent, MT = 4, many tiles per CTA
    (2, 40000, 64, 64, 3, 1, 1, 1, 1, 1),
    (1, 34368, 128, 128, 11, 1, 1, 1, 0, 0),     # the dominant layer of the B=1 step
]


def main():
    build.build(verbose=False)
    lib = _abi.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    ptr = lambda t: None if t is None else t.data_ptr()
    quick = "--quick" in sys.argv
    ok_all = True
    for case in (CASES[:6] if quick else CASES):
        B, L, Cin, Cout, K, dil, rate, use_res, acc, ragged = case
        g = torch.Generator().manual_seed(sum(case))
        x = torch.randn(B, L, Cin, generator=g)
        w = torch.randn(K, Cin, Cout, generator=g) / math.sqrt(Cin * K)
        bias = torch.randn(Cout, generator=g)
        coutR = Cout // rate
        res = torch.randn(B, L * rate, coutR, generator=g) if use_res else None
        prev = torch.randn(B, L * rate, coutR, generator=g)
        lens = torch.tensor([max(1, L // 3 - 7 * b) for b in range(B)], dtype=torch.int32, device=dev) if ragged else None
        lens_mul = 3 if ragged else 1
        for mode in (1, 0, 2, 3):
            bf = mode == 2
            if mode == 3 and Cin % 16:
                continue
            w_l = (packing.to_tc16x2_layout(w) if mode == 3 else (packing.to_tc16_layout(w) if bf else packing.to_tc_layout(w))).to(dev)
            xg = layout.to_gp(x, bf).to(dev)
            rg = layout.to_gp(res, bf).to(dev) if use_res else None
            og = layout.to_gp(prev, bf).to(dev)
            bias_d = bias.to(dev)          # keep every device tensor alive across the launch: a temporary's block would be recycled
            rc = lib.ev_op_conv1d_gp(ptr(xg), ptr(w_l), mode, ptr(bias_d), ptr(rg), ptr(og), B, L, Cin, Cout, K, dil, rate, ptr(lens), lens_mul,
                                     _abi.ACT_LRELU, 0.1, acc, 3.0, st)
            torch.cuda.synchronize()
            if rc != 0:
                print(json.dumps({"case": case, "mode": mode, "rc": rc, "err": lib.ev_last_error().decode()}), flush=True)
                ok_all = False
                continue
            got = layout.from_gp(og.cpu())                                                # (B, L*rate, coutR)
            valid = [L * rate] * B if lens is None else [min(L, int(v) * lens_mul) * rate for v in lens.tolist()]
            row = {"case": case, "mode": mode}
            if mode == 3:
                # bf16x3 fp32 emulation against an fp64 torch reference: 16 significant bits per operand -> ~1e-5 of max|ref|
                errs = []
                for b in range(B):
                    n = valid[b] // rate
                    xa = F.leaky_relu(x[b:b + 1, :n].double().transpose(1, 2), 0.1)
                    y = F.conv1d(xa, w.double().permute(2, 1, 0).contiguous(), bias.double(), padding=(K - 1) // 2 * dil, dilation=dil).transpose(1, 2)
                    y = y.reshape(1, n * rate, coutR)
                    if use_res:
                        y = y + res[b:b + 1, :n * rate].double()
                    if acc:
                        y = y + prev[b:b + 1, :n * rate].double()
                        if acc == 2:
                            y = y / 3.0
                    errs.append(float((got[b:b + 1, :n * rate].double() - y).abs().max() / y.abs().max()))
                row["rel_max_vs_fp64"] = max(errs)
                ok = max(errs) < 5e-5
            elif not bf:
                # round-1 kernel, time-major: output viewed (L, rate*coutR) == (L*rate, coutR)
                xt = x.to(dev)
                ref = prev.reshape(B, L, Cout).clone().to(dev)
                rt = res.reshape(B, L, Cout).to(dev) if use_res else None
                _abi.check(lib.ev_op_conv1d_tc(ptr(xt), ptr(w_l), 1 if mode == 1 else 0, ptr(bias_d), 0, ptr(rt), ptr(ref), B, L, Cin, Cout, K, dil,
                                               ptr(lens), lens_mul, _abi.ACT_LRELU, 0.1, _abi.ACT_NONE, acc, 3.0, None, 0, st))
                torch.cuda.synchronize()
                ref = ref.cpu().reshape(B, L * rate, coutR)
                eq = all(torch.equal(got[b, :valid[b]], ref[b, :valid[b]]) for b in range(B))
                row["bitwise_vs_tc"] = eq
                row["max_abs_diff"] = max(float((got[b, :valid[b]] - ref[b, :valid[b]]).abs().max()) for b in range(B))
                ok = eq
            else:
                # torch reference on bf16-rounded operands, fp32 accumulate
                xr = x.to(torch.bfloat16).float()
                wr = w.to(torch.bfloat16).float()
                errs = []
                for b in range(B):
                    n = valid[b] // rate
                    xa = F.leaky_relu(xr[b:b + 1, :n].transpose(1, 2), 0.1).to(torch.bfloat16).float()
                    y = F.conv1d(xa, wr.permute(2, 1, 0).contiguous(), bias, padding=(K - 1) // 2 * dil, dilation=dil).transpose(1, 2)   # (1, n, Cout)
                    y = y.reshape(1, n * rate, coutR)
                    if use_res:
                        y = y + res[b:b + 1, :n * rate].to(torch.bfloat16).float()
                    if acc:
                        y = y + prev[b:b + 1, :n * rate].to(torch.bfloat16).float()
                        if acc == 2:
                            y = y / 3.0
                    errs.append(float((got[b:b + 1, :n * rate] - y).abs().max() / y.abs().max()))
                row["rel_max_vs_torch_bf16"] = max(errs)
                ok = max(errs) < 1.2e-2          # output rounding to bf16: 2^-9 relative
            # rows >= len must be left untouched
            if lens is not None:
                pg = layout.from_gp(layout.to_gp(prev, bf))
                row["pad_rows_untouched"] = all(torch.equal(got[b, valid[b]:], pg[b, valid[b]:]) for b in range(B))
                ok = ok and row["pad_rows_untouched"]
            row["ok"] = bool(ok)
            ok_all = ok_all and ok
            print(json.dumps(row), flush=True)
    # ---- fused ResBlock layer (resblock_gp.cu) against the two conv1d_gp launches it replaces: BITWISE in every mode ----
    PAIRS = [
        # B, L, C, K, dil, acc, ragged
        (1, 300, 32, 3, 1, 0, 0),
        (2, 5000, 32, 11, 5, 0, 1),
        (3, 70000, 32, 7, 3, 1, 1),
        (2, 20000, 64, 11, 5, 2, 1),
        (2, 9000, 64, 3, 3, 0, 1),
        (1, 4000, 128, 7, 3, 0, 0),
    ]
    for case in (PAIRS[:3] if quick else PAIRS):
        B, L, C, K, dil, acc, ragged = case
        g = torch.Generator().manual_seed(sum(case) + 1)
        x = torch.randn(B, L, C, generator=g)
        w1 = torch.randn(K, C, C, generator=g) / math.sqrt(C * K)
        w2 = torch.randn(K, C, C, generator=g) / math.sqrt(C * K)
        b1, b2 = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
        prev = torch.randn(B, L, C, generator=g)
        lens = torch.tensor([max(1, L // 4 - 3 * b) for b in range(B)], dtype=torch.int32, device=dev) if ragged else None
        lens_mul = 4 if ragged else 1
        for mode in (1, 0, 2, 3):
            if mode >= 2 and C % 16:
                continue
            bf = mode == 2
            pack = packing.to_tc16x2_layout if mode == 3 else (packing.to_tc16_layout if bf else packing.to_tc_layout)
            w1d, w2d = pack(w1).to(dev), pack(w2).to(dev)
            xg = layout.to_gp(x, bf).to(dev)
            xt = torch.full_like(xg, float("nan"))
            ref = layout.to_gp(prev, bf).to(dev)
            _abi.check(lib.ev_op_conv1d_gp(ptr(xg), ptr(w1d), mode, ptr(b1), None, ptr(xt), B, L, C, C, K, dil, 1, ptr(lens), lens_mul, _abi.ACT_LRELU, 0.1,
                                           _abi.ACC_STORE, 1.0, st))
            _abi.check(lib.ev_op_conv1d_gp(ptr(xt), ptr(w2d), mode, ptr(b2), ptr(xg), ptr(ref), B, L, C, C, K, 1, 1, ptr(lens), lens_mul, _abi.ACT_LRELU, 0.1,
                                           acc, 3.0, st))
            out = layout.to_gp(prev, bf).to(dev)
            rc = lib.ev_op_resblock_gp(ptr(xg), ptr(w1d), ptr(b1), ptr(w2d), ptr(b2), mode, ptr(out), B, L, C, K, dil, ptr(lens), lens_mul, acc, 3.0, st)
            torch.cuda.synchronize()
            row = {"pair": case, "mode": mode, "rc": rc}
            if rc == 0:
                a, r = layout.from_gp(out.cpu()), layout.from_gp(ref.cpu())
                valid = [L] * B if lens is None else [min(L, int(v) * lens_mul) for v in lens.tolist()]
                row["bitwise_vs_two_launches"] = all(torch.equal(a[b, :valid[b]], r[b, :valid[b]]) for b in range(B))
                row["finite"] = all(bool(torch.isfinite(r[b, :valid[b]]).all()) for b in range(B))
                row["max_abs_diff"] = max(float((a[b, :valid[b]] - r[b, :valid[b]]).abs().max()) for b in range(B))
                pg = layout.from_gp(layout.to_gp(prev, bf))
                row["pad_rows_untouched"] = all(torch.equal(a[b, valid[b]:], pg[b, valid[b]:]) for b in range(B))
                row["ok"] = bool(row["bitwise_vs_two_launches"] and row["finite"] and row["pad_rows_untouched"])
            else:
                row["err"] = lib.ev_last_error().decode()
                row["ok"] = False
            ok_all = ok_all and row["ok"]
            print(json.dumps(row), flush=True)
    # ---- grouped launch (up to three convolutions of one shape in one kernel) against each member's own launch: BITWISE ----
    import ctypes
    GROUPS = [
        # B, L, C, Ks, dils, use_res (in place: out == res), ragged
        (1, 4296, 256, (3, 7, 11), (1, 3, 5), 0, 0),      # HiFi-GAN stage 1 at batch 1: c1 of the three ResBlocks
        (1, 4296, 256, (11, 3, 7), (1, 1, 1), 1, 0),      # ... c2, residual in place, members not sorted by taps
        (2, 3000, 128, (3, 7, 11), (5, 5, 5), 1, 1),
        (3, 900, 64, (7, 3), (3, 1), 0, 1),               # two members
        (1, 34368, 128, (3, 7, 11), (3, 3, 3), 0, 0),     # stage 2 at batch 1: 3 x 135 tiles
    ]
    for case in (GROUPS[:2] if quick else GROUPS):
        B, L, C, Ks, dils, use_res, ragged = case
        n = len(Ks)
        g = torch.Generator().manual_seed(B + L + C + sum(Ks))
        xs = [torch.randn(B, L, C, generator=g) for _ in range(n)]
        ws = [torch.randn(K, C, C, generator=g) / math.sqrt(C * K) for K in Ks]
        bs = [torch.randn(C, generator=g).to(dev) for _ in range(n)]
        prev = [torch.randn(B, L, C, generator=g) for _ in range(n)]
        lens = torch.tensor([max(1, L // 2 - 5 * b) for b in range(B)], dtype=torch.int32, device=dev) if ragged else None
        lens_mul = 2 if ragged else 1
        valid = [L] * B if lens is None else [min(L, int(v) * lens_mul) for v in lens.tolist()]
        for mode in (1, 0, 2, 3):
            bf = mode == 2
            pack = packing.to_tc16x2_layout if mode == 3 else (packing.to_tc16_layout if bf else packing.to_tc_layout)
            wd = [pack(w).to(dev) for w in ws]
            xg = [layout.to_gp(x, bf).to(dev) for x in xs]
            solo = [layout.to_gp(q, bf).to(dev) for q in prev]
            grp = [layout.to_gp(q, bf).to(dev) for q in prev]
            for i in range(n):       # with use_res the output tensor is also the residual (the engine's in-place x_j += c2(...))
                _abi.check(lib.ev_op_conv1d_gp(ptr(xg[i]), ptr(wd[i]), mode, ptr(bs[i]), ptr(solo[i]) if use_res else None, ptr(solo[i]), B, L, C, C, Ks[i], dils[i], 1,
                                               ptr(lens), lens_mul, _abi.ACT_LRELU, 0.1, _abi.ACC_STORE, 1.0, st))
            VP, FP = ctypes.c_void_p * n, ctypes.c_void_p * n
            IA = ctypes.c_int * n
            rc = lib.ev_op_conv1d_gp_group(n, VP(*[ptr(t) for t in xg]), FP(*[ptr(t) for t in wd]), mode, FP(*[ptr(t) for t in bs]),
                                           VP(*[ptr(t) for t in grp]) if use_res else None, VP(*[ptr(t) for t in grp]), IA(*Ks), IA(*dils), B, L, C, C,
                                           ptr(lens), lens_mul, _abi.ACT_LRELU, 0.1, st)
            torch.cuda.synchronize()
            row = {"group": [B, L, C, list(Ks), list(dils), use_res, ragged], "mode": mode, "rc": rc}
            if rc == 0:
                eq, pad, fin = True, True, True
                for i in range(n):
                    a, r = layout.from_gp(grp[i].cpu()), layout.from_gp(solo[i].cpu())
                    pg = layout.from_gp(layout.to_gp(prev[i], bf))
                    eq = eq and all(torch.equal(a[b, :valid[b]], r[b, :valid[b]]) for b in range(B))
                    fin = fin and all(bool(torch.isfinite(a[b, :valid[b]]).all()) for b in range(B))
                    pad = pad and all(torch.equal(a[b, valid[b]:], pg[b, valid[b]:]) for b in range(B))
                row.update(bitwise_vs_own_launches=eq, finite=fin, pad_rows_untouched=pad, ok=bool(eq and fin and pad))
            else:
                row.update(err=lib.ev_last_error().decode(), ok=False)
            ok_all = ok_all and row["ok"]
            print(json.dumps(row), flush=True)
    # ---- grouped fused ResBlock layers against each member's own ev_op_resblock_gp launch: BITWISE ----
    PGROUPS = [
        # B, L, C, Ks, dils, ragged
        (1, 68736, 64, (3, 7, 11), (1, 3, 5), 0),         # HiFi-GAN stage 3 at batch 1
        (1, 137472, 32, (11, 3, 7), (5, 5, 5), 0),        # stage 4, members not sorted by taps
        (2, 20000, 32, (3, 7, 11), (3, 3, 3), 1),
        (3, 20000, 64, (7, 3), (1, 1), 1),
    ]
    for case in (PGROUPS[:1] if quick else PGROUPS):
        B, L, C, Ks, dils, ragged = case
        n = len(Ks)
        g = torch.Generator().manual_seed(B + L + C + sum(Ks) + 7)
        xs = [torch.randn(B, L, C, generator=g) for _ in range(n)]
        w1s = [torch.randn(K, C, C, generator=g) / math.sqrt(C * K) for K in Ks]
        w2s = [torch.randn(K, C, C, generator=g) / math.sqrt(C * K) for K in Ks]
        b1s = [torch.randn(C, generator=g).to(dev) for _ in range(n)]
        b2s = [torch.randn(C, generator=g).to(dev) for _ in range(n)]
        prev = [torch.randn(B, L, C, generator=g) for _ in range(n)]
        lens = torch.tensor([max(1, L // 2 - 5 * b) for b in range(B)], dtype=torch.int32, device=dev) if ragged else None
        lens_mul = 2 if ragged else 1
        valid = [L] * B if lens is None else [min(L, int(v) * lens_mul) for v in lens.tolist()]
        for mode in (1, 0, 2, 3):
            bf = mode == 2
            pack = packing.to_tc16x2_layout if mode == 3 else (packing.to_tc16_layout if bf else packing.to_tc_layout)
            w1d, w2d = [pack(w).to(dev) for w in w1s], [pack(w).to(dev) for w in w2s]
            xg = [layout.to_gp(x, bf).to(dev) for x in xs]
            solo = [layout.to_gp(q, bf).to(dev) for q in prev]
            grp = [layout.to_gp(q, bf).to(dev) for q in prev]
            rcs = [lib.ev_op_resblock_gp(ptr(xg[i]), ptr(w1d[i]), ptr(b1s[i]), ptr(w2d[i]), ptr(b2s[i]), mode, ptr(solo[i]), B, L, C, Ks[i], dils[i], ptr(lens), lens_mul,
                                         _abi.ACC_STORE, 1.0, st) for i in range(n)]
            VP, IA = ctypes.c_void_p * n, ctypes.c_int * n
            tab = lambda ts: VP(*[ptr(t) for t in ts])
            rc = lib.ev_op_resblock_gp_group(n, tab(xg), tab(w1d), tab(b1s), tab(w2d), tab(b2s), mode, tab(grp), B, L, C, IA(*Ks), IA(*dils), ptr(lens), lens_mul, st)
            torch.cuda.synchronize()
            row = {"pair_group": [B, L, C, list(Ks), list(dils), ragged], "mode": mode, "rc": rc, "solo_rc": rcs}
            # a member whose own plan keeps fewer than two accumulators per tile (3xTF32 at 64 channels) is not grouped: EV_EINVAL expected
            v11 = (ctypes.c_int * 11)()
            mts = [v11[0] if lib.ev_debug_resblock_gp_plan(B, L, C, Ks[i], dils[i], mode, v11) == 0 else 0 for i in range(n)]
            if min(mts) < 2:
                row.update(member_mt=mts, ok=bool(rc != 0))
            elif rc == 0 and not any(rcs):
                eq, pad, fin = True, True, True
                for i in range(n):
                    a, r = layout.from_gp(grp[i].cpu()), layout.from_gp(solo[i].cpu())
                    pg = layout.from_gp(layout.to_gp(prev[i], bf))
                    eq = eq and all(torch.equal(a[b, :valid[b]], r[b, :valid[b]]) for b in range(B))
                    fin = fin and all(bool(torch.isfinite(a[b, :valid[b]]).all()) for b in range(B))
                    pad = pad and all(torch.equal(a[b, valid[b]:], pg[b, valid[b]:]) for b in range(B))
                row.update(bitwise_vs_own_launches=eq, finite=fin, pad_rows_untouched=pad, ok=bool(eq and fin and pad))
            else:
                row.update(err=lib.ev_last_error().decode(), ok=False)
            ok_all = ok_all and row["ok"]
            print(json.dumps(row), flush=True)
    # boundary kernels
    g = torch.Generator().manual_seed(5)
    mel = torch.randn(2, 80, 37, generator=g)
    mel_d = mel.to(dev)
    for bf in (0, 1):
        outg = torch.empty((2, 80 // (8 if bf else 4), 37, 8 if bf else 4), dtype=torch.bfloat16 if bf else torch.float32, device=dev)
        _abi.check(lib.ev_op_to_gp(ptr(mel_d), 80 * 37, 1, 37, ptr(outg), 2, 37, 80, bf, st))
        torch.cuda.synchronize()
        eq = torch.equal(outg.cpu(), layout.to_gp(mel.transpose(1, 2).contiguous(), bool(bf)))
        print(json.dumps({"to_gp_channels_first": eq, "bf16": bf}), flush=True)
        ok_all = ok_all and eq
    x = torch.randn(2, 5000, 32, generator=g) * 2
    w = torch.randn(7, 32, generator=g) * 0.1
    b1 = torch.randn(1, generator=g)
    lens = torch.tensor([20, 13], dtype=torch.int32, device=dev)
    w_d, b1_d = w.to(dev), b1.to(dev)
    # the time-major conv_post is not exported as an op: compare with torch
    ref = torch.tanh(F.conv1d(F.leaky_relu(x, 0.01).transpose(1, 2), w.t().unsqueeze(0), b1, padding=3))[:, 0]
    for bf in (0, 1):
        xg = layout.to_gp(x, bool(bf)).to(dev)
        wav = torch.empty(2, 5000, device=dev)
        _abi.check(lib.ev_op_conv_post_gp(ptr(xg), bf, ptr(w_d), ptr(b1_d), ptr(lens), 256, 2, 5000, 32, 7, 0.01, ptr(wav), st))
        torch.cuda.synchronize()
        e = max(float((wav[b, :int(lens[b]) * 256 - 3].cpu() - ref[b, :int(lens[b]) * 256 - 3]).abs().max()) for b in range(2))
        z = bool((wav[1, 13 * 256:] == 0).all())
        okp = e < (2e-2 if bf else 1e-5) and z
        print(json.dumps({"conv_post_gp_max_abs_err": e, "pad_zero": z, "bf16": bf, "ok": okp}), flush=True)
        ok_all = ok_all and okp
    # odd length (scalar stores), no lengths, the generic-K kernel (K = 5) beside the K = 7 one
    for K in (7, 5):
        x = torch.randn(1, 1001, 32, generator=g)
        w = torch.randn(K, 32, generator=g) * 0.1
        ref = torch.tanh(F.conv1d(F.leaky_relu(x, 0.01).transpose(1, 2), w.t().unsqueeze(0), b1, padding=K // 2))[:, 0]
        xg, w_d = layout.to_gp(x, False).to(dev), w.to(dev)
        wav = torch.empty(1, 1001, device=dev)
        _abi.check(lib.ev_op_conv_post_gp(ptr(xg), 0, ptr(w_d), ptr(b1_d), None, 1, 1, 1001, 32, K, 0.01, ptr(wav), st))
        torch.cuda.synchronize()
        e = float((wav.cpu() - ref).abs().max())
        print(json.dumps({"conv_post_gp_odd_len_K": K, "max_abs_err": e, "ok": e < 1e-5}), flush=True)
        ok_all = ok_all and e < 1e-5
    print("GP_CHECK_" + ("OK" if ok_all else "FAILED"), flush=True)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
