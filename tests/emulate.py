"""CPU emulation helpers for tests (test infrastructure only, never used by the product path).

  emulate_packed_conv : executes the kernel's MMA schedule (K blocks x taps x 16-channel k-steps with
                        shifted halo rows) from the packed weights returned by ss4k_debug_pack;
                        checks the weight packer / schedule tables of csrc/engine.cu without a GPU.
  run_program         : interprets a layer program (ss4k_plan_dry JSON) with torch fp32 ops; checks the
                        network lowering of csrc/program.cpp (buffers, channel offsets, epilogues).
"""
import ctypes
import json

import torch
import torch.nn.functional as F


def debug_pack(lib, L, w, bias=None, slope=None, mode=0, in_pitch=None, in_coff=0, wperm=0, act_mode=0,
               n=1, h=8, wd=8, desc_mode=0):
    cout, cin = w.shape[:2]
    d = L.ConvDesc()
    d.struct_size = ctypes.sizeof(L.ConvDesc)
    d.n, d.h, d.w, d.cin, d.cout, d.mode, d.act_mode = n, h, wd, cin, cout, mode, act_mode
    d.reserved[1] = desc_mode
    if in_pitch is None:
        in_pitch = (cin + 15) // 16 * 16
    w = w.contiguous().float()
    js, pk, cnt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    bp = ctypes.c_void_p(bias.contiguous().float().data_ptr()) if bias is not None else None
    sp = ctypes.c_void_p(slope.contiguous().float().data_ptr()) if slope is not None else None
    L.check(lib.ss4k_debug_pack(ctypes.byref(d), in_pitch, in_coff, wperm, ctypes.c_void_p(w.data_ptr()), bp, sp,
                                ctypes.byref(js), ctypes.byref(pk), ctypes.byref(cnt)))
    meta = json.loads(ctypes.string_at(js).decode())
    arr = (ctypes.c_float * cnt.value).from_address(pk.value)
    packed = torch.tensor(list(arr), dtype=torch.float32).reshape(meta["nkb"], meta["ntaps"], meta["npad"], 64)
    lib.ss4k_free(js)
    lib.ss4k_free(pk)
    return meta, packed


def emulate_packed_conv(meta, packed, x_nhwc, mode, x_lo=None):
    """x_nhwc: [N,H,W,pitch] float.  Returns the accumulators in OUTPUT pixel space [N,OH,OW,npad]
    (before bias / activation), computed exactly the way the kernel schedules its MMAs."""
    n, h, w, pitch = x_nhwc.shape
    srcs = [x_nhwc, x_lo if x_lo is not None else x_nhwc]
    if mode == 2:   # stride-2 view: (2*pitch merged (pb,c), W/2, pa, H/2)
        ah, aw = h // 2, w // 2
        def plane(t, p):
            v = t.reshape(n, ah, 2, aw, 2, pitch)[:, :, p]          # [n, ah, aw, pb, c]
            return v.reshape(n, ah, aw, 2 * pitch)
    else:
        ah, aw = h, w
        def plane(t, p):
            return t
    nsub, npad = meta["nsub"], meta["npad"]
    acc = torch.zeros(nsub, n, ah, aw, npad, dtype=torch.float64)
    for kbi, (tmap, c0, p) in enumerate(meta["kb"]):
        src = plane(srcs[tmap], p)
        cdim = src.shape[-1]
        blk = torch.zeros(n, ah, aw, 64, dtype=torch.float64)
        hi = min(c0 + 64, cdim)
        if hi > c0:
            blk[..., :hi - c0] = src[..., c0:hi]                     # TMA zero-fills beyond the tensor
        padded = F.pad(blk.permute(0, 3, 1, 2), (2, 2, 2, 2)).permute(0, 2, 3, 1)   # zero halo
        for t, (dr, shift, sub) in enumerate(meta["taps"]):
            mask = meta["mask"][kbi][t]
            if not mask:
                continue
            # output (y, x) reads input row y - 1 + dr, pixel x - 1 + shift
            win = padded[:, 1 + dr:1 + dr + ah, 1 + shift:1 + shift + aw, :]
            for ks in range(4):
                if mask >> ks & 1:
                    a = win[..., ks * 16:(ks + 1) * 16]
                    b = packed[kbi, t, :, ks * 16:(ks + 1) * 16].double()
                    acc[sub] += a @ b.t()
    if mode == 1:   # 4 output phases -> 2x resolution
        out = torch.zeros(n, 2 * ah, 2 * aw, npad, dtype=torch.float64)
        for sub in range(4):
            out[:, (sub >> 1)::2, (sub & 1)::2] = acc[sub]
        return out
    return acc[0]


# ------------------------------------------------------------------------------------------------
def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def run_program(prog, sd, x, quant=None):
    """Interpret a layer program with fp32 torch ops.  x: the caller's input (float NCHW here).
    quant: optional callable emulating 16-bit storage of activations (e.g. lambda t: t.half().float())."""
    q = quant if quant is not None else (lambda t: t)
    bufs = [torch.zeros(b["n"], b["h"], b["w"], b["pitch"]) for b in prog["bufs"]]
    result = None
    for st in prog["steps"]:
        if st["kind"] == "prep":
            assert st["in_fmt"] == 0
            us = st["unshuffle"]
            src = x
            if x.ndim == 5:   # BSVD: [N, F, C, H, W] is one stream of N*F frames (model.py:519-520)
                src = x.reshape(-1, *x.shape[2:])
            if us > 1:
                src = F.pixel_unshuffle(x, us)
            n0, nn = st.get("n0", 0), st["n"]          # frames [n0, n0 + n) of the clip (BSVD chunk with a temporal halo)
            dst = bufs[st["out_buf"]]
            dst.zero_()
            dst[n0:n0 + nn, ..., :src.shape[1]] = q(_nhwc(src[n0:n0 + nn]))
            if st["fill_ch"] >= 0:
                dst[n0:n0 + nn, ..., st["fill_ch"]] = q(torch.tensor(st["fill_val"]))
            continue
        c = st
        cin, cout = c["cin"], c["cout"]
        n0, nn = c.get("n0", 0), c["n"]
        fr = slice(n0, n0 + nn)                          # the frames this step processes
        src = _nchw(bufs[c["in_buf"]][fr][..., c["in_coff"]:c["in_coff"] + cin])
        w, b = sd[c["wname"]], sd[c["bname"]]
        if c.get("neg_first", 0):
            w, b = w.clone(), b.clone()
            w[:c["neg_first"]] *= -1
            b[:c["neg_first"]] *= -1
        if c["mode"] == 0:
            v = F.conv2d(src, w, b, padding=1)
        elif c["mode"] == 1:
            v = F.conv2d(F.interpolate(src, scale_factor=2, mode="nearest"), w, b, padding=1)
        else:
            v = F.conv2d(src, w, b, stride=2, padding=1)
        if c["act"] == 1:
            slope = sd[c["sname"]].view(1, -1, 1, 1) if c["sname"] else torch.full((1, cout, 1, 1), c["const_slope"])
            v = torch.where(v >= 0, v, v * slope)
        elif c["act"] == 2:
            v = torch.clamp(v, 0, 6)
        v = v * c["alpha"]
        om = c["out_mode"]
        if om == 3:      # PixelShuffle(2) into NHWC
            v = F.pixel_shuffle(v, 2)
        oc = v.shape[1]
        for k in (1, 2):
            rb = c[f"res{k}_buf"]
            if rb >= 0:
                if k == 1 and c.get("res1_nch", 0):
                    nch = c["res1_nch"]
                    r = torch.zeros_like(v)
                    r[:, :nch] = _nchw(bufs[rb][fr][..., c["res1_coff"]:c["res1_coff"] + nch])
                else:
                    r = _nchw(bufs[rb][fr][..., c[f"res{k}_coff"]:c[f"res{k}_coff"] + oc])
                v = v + c[f"beta{k}"] * r
        if om in (0, 3):
            dst = bufs[c["out_buf"]]
            if c.get("up2_store", 0):   # nearest-x2 upsample fused into the store
                assert (c["out_h"], c["out_w"]) == tuple(v.shape[2:])
                v = F.interpolate(v, scale_factor=2, mode="nearest")
                assert tuple(dst.shape[1:3]) == tuple(v.shape[2:])
            else:
                assert tuple(dst.shape[1:3]) == (c["out_h"], c["out_w"]) == tuple(v.shape[2:]), (c["name"], dst.shape, v.shape)
            npad = (oc + 15) // 16 * 16
            vv = q(_nhwc(v))
            if c.get("tshift", 0):   # temporal-shift scatter: time == batch index, out-of-clip slices dropped
                fold, o = c["fold"], c["out_coff"]
                T = dst.shape[0]
                for i in range(nn):
                    t = n0 + i
                    if t > 0:
                        dst[t - 1, ..., o:o + fold] = vv[i, ..., :fold]
                    if t < T - 1:
                        dst[t + 1, ..., o + fold:o + 2 * fold] = vv[i, ..., fold:2 * fold]
                dst[fr][..., o + 2 * fold:o + oc] = vv[..., 2 * fold:]
            else:
                dst[fr][..., c["out_coff"]:c["out_coff"] + npad] = 0
                dst[fr][..., c["out_coff"]:c["out_coff"] + oc] = vv
        elif om == 4:    # temporal-shift scatter
            fold = c["fold"]
            vv = q(_nhwc(v))
            bufs[c["out2_buf"]][..., c["out_coff"]:c["out_coff"] + fold] = vv[..., :fold]
            bufs[c["out3_buf"]][..., c["out_coff"] + fold:c["out_coff"] + 2 * fold] = vv[..., fold:2 * fold]
            bufs[c["out_buf"]][..., c["out_coff"] + 2 * fold:c["out_coff"] + oc] = vv[..., 2 * fold:]
        elif om in (1, 5):
            result = v if om == 1 else v.half().float()
        elif om == 6:
            f = torch.clamp(v, 0, 1) * 255
            result = (torch.round(f) if c["round_u8"] else f).to(torch.uint8).permute(0, 2, 3, 1)
        elif om in (2, 7):    # PixelShuffle(r) + nearest-upsampled base image
            r = c["ps_r"]
            v = F.pixel_shuffle(v, r)
            base = _nchw(bufs[c["base_buf"]][..., :v.shape[1]])
            result = v + F.interpolate(base, scale_factor=float(r), mode="nearest")
        else:
            raise AssertionError(om)
    return result


# ------------------------------------------------------------------------------------------------
def debug_pack_stream(lib, L, w, bias=None, slope=None, in_pitch=None, in_coff=0, wperm=0, act_mode=0, act=0, alpha=1.0,
                      mode=0):
    """Weight layout + configuration of the row-streaming kernel (csrc/conv_stream.cu)."""
    cout, cin = w.shape[:2]
    d = L.ConvDesc()
    d.struct_size = ctypes.sizeof(L.ConvDesc)
    d.n, d.h, d.w, d.cin, d.cout, d.mode, d.act_mode = 1, 8, 8, cin, cout, mode, act_mode
    d.act, d.alpha = act, alpha
    d.reserved[6] = 1
    if in_pitch is None:
        in_pitch = (cin + 15) // 16 * 16
    w = w.contiguous().float()
    js, pk, cnt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    bp = ctypes.c_void_p(bias.contiguous().float().data_ptr()) if bias is not None else None
    sp = ctypes.c_void_p(slope.contiguous().float().data_ptr()) if slope is not None else None
    L.check(lib.ss4k_debug_pack(ctypes.byref(d), in_pitch, in_coff, wperm, ctypes.c_void_p(w.data_ptr()), bp, sp,
                                ctypes.byref(js), ctypes.byref(pk), ctypes.byref(cnt)))
    meta = json.loads(ctypes.string_at(js).decode())
    arr = (ctypes.c_float * cnt.value).from_address(pk.value)
    flat = torch.tensor(list(arr), dtype=torch.float32).reshape(-1, 64)
    # weight tile groups: one per K block, except that the A_hi*W_hi and A_lo*W_hi K blocks of the split mode share one
    packed = flat[:meta["w_rows"]].reshape(meta["chunks"], meta["nwt"], meta["nkx"], 3, meta["nout"], 64)[:, meta["wt"]]
    # (the 64-wide variant appends one bias tile per chunk: bias hi / lo halves in K columns 0 / 1, for its bias MMA)
    meta["bias_tiles"] = flat[meta["w_rows"]:]
    # meta["bias_f"]: fp32 bias with alpha folded in, [chunks * nout] -- the accumulators' initial value
    lib.ss4k_free(js)
    lib.ss4k_free(pk)
    return meta, packed


def emulate_stream_conv(meta, packed, x_nhwc, grid, in_coff=0, acc_slots=None):
    """Mirror of conv3x3_stream_kernel's schedule: per-CTA bands of output rows, accumulator ring with
    vertically fused taps (one 'MMA' adds input row r into the slots of output rows r-1, r, r+1, split
    where the ring wraps), completion commits, in-order epilogue drain.  Fresh slots hold the bias row
    (written by the epilogue warps after the drain), every MMA accumulates.  Returns the accumulators [N, H, W, npad]
    (alpha-folded conv + bias, before the activation)."""
    n_img, H, W, pitch = x_nhwc.shape
    nout, chunks, nkb, S = meta["nout"], meta["chunks"], meta["nkb"], acc_slots or meta["acc_slots"]
    strips = (W + 127) // 128
    total = chunks * n_img * strips * H
    out = torch.full((n_img, H, W, chunks * nout), float("nan"), dtype=torch.float64)
    xpad = torch.zeros(n_img, H, strips * 128 + 2, (in_coff // 64 + nkb) * 64 + 64, dtype=torch.float64)
    xpad[:, :, 1:W + 1, :pitch] = x_nhwc.double()   # column -1 / >= W read zeros (TMA OOB fill)
    g = min(grid, total)
    for cta in range(g):
        u, u1 = cta * total // g, (cta + 1) * total // g
        tmem = torch.zeros(S, 128, nout, dtype=torch.float64)
        state = ["empty"] * S          # empty -> busy (being accumulated) -> full (committed) -> empty (drained)
        qs = 0
        while u < u1:
            t = u
            y = t % H; t //= H
            strip = t % strips; t //= strips
            n = t % n_img; chunk = t // n_img
            yb, ye = y, min(H, y + (u1 - u))
            u += ye - yb
            r0, r1 = max(yb - 1, 0), min(ye, H - 1)
            for r in range(r0, r1 + 1):
                y_lo, y_hi = max(r - 1, yb), min(r + 1, ye - 1)
                b_lo, b_hi = y_lo - (r - 1), y_hi - (r - 1)
                f_lo = y_lo if r == r0 else r + 1
                bias_f = torch.tensor(meta["bias_f"], dtype=torch.float64)
                for yy in range(f_lo, y_hi + 1):
                    s = (qs + yy - yb) % S
                    assert state[s] == "empty", ("accumulator slot not drained", cta, r, yy, s, state)
                    state[s] = "busy"
                    tmem[s] = bias_f[chunk * nout:(chunk + 1) * nout].expand(128, nout).clone()   # tcgen05.st by the epilogue

                s0 = (qs + (y_lo - yb)) % S
                nblk = b_hi - b_lo + 1
                nA = nblk if s0 + nblk <= S else S - s0
                ops = [(s0, b_lo, nA)] + ([(0, b_lo + nA, nblk - nA)] if nblk > nA else [])
                for kb in range(nkb):
                    c0 = in_coff + kb * 64
                    slab = xpad[n, r, strip * 128:strip * 128 + 130, c0:c0 + 64]      # 130-pixel halo row
                    for kx in range(3):
                        a_full = slab[kx:kx + 128]                                    # shifted start address
                        for ks in range(meta["nks"][kb]):
                            a = a_full[:, ks * 16:(ks + 1) * 16]
                            for (sa, b0, nb_) in ops:
                                wt = packed[chunk, kb, kx, b0:b0 + nb_, :, ks * 16:(ks + 1) * 16].double()  # [nb_, nout, 16]
                                d = torch.einsum("mk,bnk->bmn", a, wt)
                                for i in range(nb_):
                                    assert state[sa + i] == "busy"
                                    tmem[sa + i] += d[i]
                done = []
                if r - 1 >= yb:
                    done.append(r - 1)
                if r == r1 and r <= ye - 1:
                    done.append(r)
                for yy in done:           # commit -> epilogue drains (in order) -> slot free
                    s = (qs + yy - yb) % S
                    assert state[s] == "busy"
                    x0 = strip * 128
                    wv = min(128, W - x0)
                    out[n, yy, x0:x0 + wv, chunk * nout:(chunk + 1) * nout] = tmem[s, :wv]
                    state[s] = "empty"
            qs = (qs + ye - yb) % S
        assert all(st == "empty" for st in state)
    assert not torch.isnan(out).any()
    return out


def emulate_stream_conv_s2(meta, packed, x_nhwc, grid, acc_slots=None):
    """Mirror of conv3x3_stream_kernel's STRIDE-2 schedule (StreamParams::stride2): the input is read through its
    pixel-pair view (2*pitch channels per pair), two horizontal shifts with k-step masks, N blocks [ky2 | ky0 | ky1]:
    an odd input row feeds output rows (r-1)/2 and (r+1)/2 with one 'MMA', an even row feeds row r/2.
    Returns the accumulators [N, H/2, W/2, npad]."""
    n_img, Hin, Win, pitch = x_nhwc.shape
    H, W = Hin // 2, Win // 2
    nout, chunks, nkb, S = meta["nout"], meta["chunks"], meta["nkb"], acc_slots or meta["acc_slots"]
    assert meta["stride2"] == 1 and meta["nkx"] == 2
    strips = (W + 127) // 128
    total = chunks * n_img * strips * H
    out = torch.full((n_img, H, W, chunks * nout), float("nan"), dtype=torch.float64)
    nblk_src = 2 * pitch // 64
    pairs = x_nhwc.double().reshape(n_img, Hin, W, 2 * pitch)
    xpad = torch.zeros(n_img, Hin, strips * 128 + 2, nblk_src * 64, dtype=torch.float64)
    xpad[:, :, 1:W + 1] = pairs                      # pair -1 / >= W read zeros (TMA OOB fill)
    bias_f = torch.tensor(meta["bias_f"], dtype=torch.float64)
    g = min(grid, total)
    for cta in range(g):
        u, u1 = cta * total // g, (cta + 1) * total // g
        tmem = torch.zeros(S, 128, nout, dtype=torch.float64)
        state = ["empty"] * S
        qs = 0
        while u < u1:
            t = u
            y = t % H; t //= H
            strip = t % strips; t //= strips
            n = t % n_img; chunk = t // n_img
            yb, ye = y, min(H, y + (u1 - u))
            u += ye - yb
            r0, r1 = (2 * yb - 1 if yb > 0 else 0), 2 * ye - 1
            for r in range(r0, r1 + 1):
                if r & 1:
                    ya = (r - 1) // 2
                    y_lo, y_hi = max(ya, yb), min(ya + 1, ye - 1)
                    b_lo, f_lo, done = y_lo - ya, ya + 1, ([ya] if ya >= yb else [])
                else:
                    y_lo = y_hi = r // 2
                    b_lo, f_lo, done = 2, y_hi + 1, []
                if r == r0:
                    f_lo = y_lo
                for yy in range(f_lo, y_hi + 1):
                    s = (qs + yy - yb) % S
                    assert state[s] == "empty", ("accumulator slot not drained", cta, r, yy, s, state)
                    state[s] = "busy"
                    tmem[s] = bias_f[chunk * nout:(chunk + 1) * nout].expand(128, nout).clone()
                s0 = (qs + (y_lo - yb)) % S
                nblk = y_hi - y_lo + 1
                nA = nblk if s0 + nblk <= S else S - s0
                ops = [(s0, b_lo, nA)] + ([(0, b_lo + nA, nblk - nA)] if nblk > nA else [])
                for kb in range(nkb):
                    c0 = meta["src_kb"][kb] * 64
                    slab = xpad[n, r, strip * 128:strip * 128 + 130, c0:c0 + 64]
                    for sh in range(2):
                        a_full = slab[sh:sh + 128]
                        for ks in range(4):
                            if not (meta["ksm"][kb][sh] >> ks) & 1:
                                continue
                            a = a_full[:, ks * 16:(ks + 1) * 16]
                            for (sa, b0, nb_) in ops:
                                wt = packed[chunk, kb, sh, b0:b0 + nb_, :, ks * 16:(ks + 1) * 16].double()
                                d = torch.einsum("mk,bnk->bmn", a, wt)
                                for i in range(nb_):
                                    assert state[sa + i] == "busy"
                                    tmem[sa + i] += d[i]
                for yy in done:
                    s = (qs + yy - yb) % S
                    assert state[s] == "busy"
                    x0 = strip * 128
                    wv = min(128, W - x0)
                    out[n, yy, x0:x0 + wv, chunk * nout:(chunk + 1) * nout] = tmem[s, :wv]
                    state[s] = "empty"
            qs = (qs + ye - yb) % S
        assert all(st == "empty" for st in state)
    assert not torch.isnan(out).any()
    return out
