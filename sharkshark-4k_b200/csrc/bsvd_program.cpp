// BSVD clip program: the reference's streaming denoiser restated over a whole clip, time == batch index.
//
// Reference (all paths relative to the reference tree):
//   BSVD.forward / streaming_forward        src/upscale/model/bsvd/model.py:515-580
//   DenBlock.forward                         model.py:402-424  (skip FIFOs :332-350, none_minus :436-442)
//   BiBufferConv / ShiftConv                 model.py:22-138   (fold = C/8; X_{t+1}[:fold], X_{t-1}[fold:2fold], X_t[2fold:])
//   service config                           src/upscale/model/bsvd/factory.py:31-35 (chns 32/64/128, mid 32, interm 30, relu6)
//
// Every BiBufferConv delays the stream by one frame, so the streaming network equals a frame-aligned
// network over the clip in which a shift conv reads its three channel slices from frames t+1, t-1, t
// (zero features outside the clip).  The temporal gather is done on the PRODUCER side: the conv that
// feeds a shift conv stores channels [0, fold) of frame t into frame t-1's tensor and channels
// [fold, 2 fold) into frame t+1's, so the consumer's TMA loader reads one contiguous pixel row.
#include "program.h"

#include <algorithm>
#include <vector>

#include "conv_params.h"

namespace ss4k {

namespace {
int round_up16(int a) { return (a + 15) / 16 * 16; }
}  // namespace

std::string build_bsvd_clip(const PlanCfgLite& c, Program* P) {
  const int T = c.n, H = c.h, W = c.w;
  if (H % 4 || W % 4) return "BSVD: H and W must be multiples of 4 (two stride-2 stages)";
  if (c.out_fmt != 0 && c.out_fmt != 1) return "BSVD: output is float / half NCHW";
  const int c0 = 32, c1 = 64, c2 = 128, mid = 32, interm = 30;
  // fp16 hi/lo split precision (SS4K_ACT_F16_SPLIT): every activation tensor has a low-half twin, every product is
  // three MMAs (hi*hi + hi*lo + lo*hi).  Needed for parity with ill-conditioned weights such as the reference
  // constructor's kaiming init (fp32 outputs span +-14); clip and streaming layouts alike
  // (a twin is one more ring of frames with the same geometry, bsvd_stream.inc).
  const bool split = c.act_mode == 2;
  P->in_n = T; P->in_c = 4; P->in_h = H; P->in_w = W;
  P->out_n = T; P->out_c = 3; P->out_h = H; P->out_w = W;
  const int in16 = P->add_buf("in16", T, H, W, 16);
  const int x0a = P->add_buf("x0a", T, H, W, 32);
  const int x0 = P->add_buf("x0", T, H, W, c0);
  const int d0 = P->add_buf("d0", T, H / 2, W / 2, c1, true);
  const int m0a = P->add_buf("m0a", T, H / 2, W / 2, c1, true);
  const int x1 = P->add_buf("x1", T, H / 2, W / 2, c1);
  const int d1 = P->add_buf("d1", T, H / 4, W / 4, c2, true);
  const int m1a = P->add_buf("m1a", T, H / 4, W / 4, c2, true);
  const int m1b = P->add_buf("m1b", T, H / 4, W / 4, c2, true);
  const int u2a = P->add_buf("u2a", T, H / 4, W / 4, c2, true);
  const int u2b = P->add_buf("u2b", T, H / 4, W / 4, c2);
  const int p2 = P->add_buf("p2", T, H / 2, W / 2, c1, true);
  const int u1a = P->add_buf("u1a", T, H / 2, W / 2, c1, true);
  const int u1b = P->add_buf("u1b", T, H / 2, W / 2, c1);
  const int p1 = P->add_buf("p1", T, H, W, c0);
  const int o0 = P->add_buf("o0", T, H, W, c0);
  const int t1out = P->add_buf("t1out", T, H, W, mid);
  // streaming: both DenBlocks are in flight at once (temp2 runs 8 frames behind temp1) -> own buffers
  int alt[16];
  const int firsts[16] = {x0a, x0, d0, m0a, x1, d1, m1a, m1b, u2a, u2b, p2, u1a, u1b, p1, o0, -1};
  for (int i = 0; i < 15; ++i) {
    alt[i] = firsts[i];
    if (c.bsvd_stream) {
      const BufSpec bs = P->bufs[firsts[i]];
      alt[i] = P->add_buf(bs.name + "_2", bs.n, bs.h, bs.w, bs.pitch, bs.zero_init);
    }
  }

  std::vector<int> lo_of;  // buffer id -> id of its low-half twin
  if (split) {
    const int nb = static_cast<int>(P->bufs.size());
    lo_of.assign(nb, kBufNone);
    for (int i = 0; i < nb; ++i) {
      const BufSpec bs = P->bufs[i];
      lo_of[i] = P->add_buf(bs.name + "_lo", bs.n, bs.h, bs.w, bs.pitch, bs.zero_init);
    }
  }
  auto lo = [&](int id) { return split && id >= 0 && id < static_cast<int>(lo_of.size()) ? lo_of[id] : kBufNone; };

  PrepSpec pp;
  pp.in_fmt = c.in_fmt; pp.c = 4; pp.h = H; pp.w = W; pp.n = T; pp.out_buf = in16; pp.out_lo_buf = lo(in16);
  if (c.in_fmt == 2 || c.in_fmt == 3) {  // uint8 NHWC RGB / NV12 frames: the noise map (fsrcnn_upscaler.py:262) is filled in
    pp.c = 3; pp.fill_ch = 3; pp.fill_val = c.bsvd_noise;
    P->in_c = 3;
  }
  P->add_prep(pp);

  auto conv = [&](const std::string& nm, int mode, int in_buf, int ih, int iw, int ipitch, int cin, int cout, int act) {
    ConvSpec v;
    v.name = nm; v.wname = nm + ".weight"; v.bname = nm + ".bias";
    v.mode = mode; v.n = T; v.cin = cin; v.cout = cout;
    v.in_buf = in_buf; v.in_h = ih; v.in_w = iw; v.in_pitch = ipitch;
    v.in_lo_buf = lo(in_buf); v.split = split ? 1 : 0;
    v.act = act;
    v.out_mode = kOutNHWC;
    v.out_h = mode == kModeS2 ? ih / 2 : ih;
    v.out_w = mode == kModeS2 ? iw / 2 : iw;
    return v;
  };
  auto shifted = [&](ConvSpec& v, int out_buf, int C) {  // the consumer is a shift conv
    v.out_buf = out_buf; v.out_lo_buf = lo(out_buf); v.out_pitch = C; v.tshift = 1; v.fold = C / 8;
  };
  auto plain = [&](ConvSpec& v, int out_buf, int pitch) { v.out_buf = out_buf; v.out_lo_buf = lo(out_buf); v.out_pitch = pitch; };

  auto den_block = [&](const std::string& p, int in_buf, int in_pitch, int in_c, int out_c, bool last) {
    // (temp2 of the streaming layout uses the second buffer set)
    const int* B = last ? alt : firsts;
    const int x0a = B[0], x0 = B[1], d0 = B[2], m0a = B[3], x1 = B[4], d1 = B[5], m1a = B[6], m1b = B[7], u2a = B[8],
              u2b = B[9], p2 = B[10], u1a = B[11], u1b = B[12], p1 = B[13], o0 = B[14];
    { ConvSpec v = conv(p + "inc.convblock.0", kModeConv3, in_buf, H, W, in_pitch, in_c, interm, kActRelu6); plain(v, x0a, 32); P->add_conv(v); }
    { ConvSpec v = conv(p + "inc.convblock.3", kModeConv3, x0a, H, W, 32, interm, c0, kActRelu6); plain(v, x0, c0); P->add_conv(v); }
    { ConvSpec v = conv(p + "downc0.convblock.0", kModeS2, x0, H, W, c0, c0, c1, kActRelu6); shifted(v, d0, c1); P->add_conv(v); }
    { ConvSpec v = conv(p + "downc0.memconv.c1.op.conv", kModeConv3, d0, H / 2, W / 2, c1, c1, c1, kActRelu6); shifted(v, m0a, c1); P->add_conv(v); }
    { ConvSpec v = conv(p + "downc0.memconv.c2.op.conv", kModeConv3, m0a, H / 2, W / 2, c1, c1, c1, kActRelu6); plain(v, x1, c1); P->add_conv(v); }
    { ConvSpec v = conv(p + "downc1.convblock.0", kModeS2, x1, H / 2, W / 2, c1, c1, c2, kActRelu6); shifted(v, d1, c2); P->add_conv(v); }
    { ConvSpec v = conv(p + "downc1.memconv.c1.op.conv", kModeConv3, d1, H / 4, W / 4, c2, c2, c2, kActRelu6); shifted(v, m1a, c2); P->add_conv(v); }
    { ConvSpec v = conv(p + "downc1.memconv.c2.op.conv", kModeConv3, m1a, H / 4, W / 4, c2, c2, c2, kActRelu6); shifted(v, m1b, c2); P->add_conv(v); }
    { ConvSpec v = conv(p + "upc2.memconv.c1.op.conv", kModeConv3, m1b, H / 4, W / 4, c2, c2, c2, kActRelu6); shifted(v, u2a, c2); P->add_conv(v); }
    { ConvSpec v = conv(p + "upc2.memconv.c2.op.conv", kModeConv3, u2a, H / 4, W / 4, c2, c2, c2, kActRelu6); plain(v, u2b, c2); P->add_conv(v); }
    {  // conv + PixelShuffle(2) + skip3, feeding upc1's first shift conv
      ConvSpec v = conv(p + "upc2.convblock.0", kModeConv3, u2b, H / 4, W / 4, c2, c2, c1 * 4, kActNone);
      v.out_mode = kOutPS2NHWC; v.wperm = 1; v.out_h = H / 2; v.out_w = W / 2;
      v.res1_buf = x1; v.res1_lo_buf = lo(x1); v.res1_pitch = c1; v.beta1 = 1.f;
      shifted(v, p2, c1);
      P->add_conv(v);
    }
    { ConvSpec v = conv(p + "upc1.memconv.c1.op.conv", kModeConv3, p2, H / 2, W / 2, c1, c1, c1, kActRelu6); shifted(v, u1a, c1); P->add_conv(v); }
    { ConvSpec v = conv(p + "upc1.memconv.c2.op.conv", kModeConv3, u1a, H / 2, W / 2, c1, c1, c1, kActRelu6); plain(v, u1b, c1); P->add_conv(v); }
    {  // conv + PixelShuffle(2) + skip2
      ConvSpec v = conv(p + "upc1.convblock.0", kModeConv3, u1b, H / 2, W / 2, c1, c1, c0 * 4, kActNone);
      v.out_mode = kOutPS2NHWC; v.wperm = 1; v.out_h = H; v.out_w = W;
      v.res1_buf = x0; v.res1_lo_buf = lo(x0); v.res1_pitch = c0; v.beta1 = 1.f;
      plain(v, p1, c0);
      P->add_conv(v);
    }
    { ConvSpec v = conv(p + "outc.convblock.0", kModeConv3, p1, H, W, c0, c0, c0, kActRelu6); plain(v, o0, c0); P->add_conv(v); }
    {  // out[:, :3] = in[:, :3] - out[:, :3]: first three output channels negated in the weights, + masked input residual
      ConvSpec v = conv(p + "outc.convblock.3", kModeConv3, o0, H, W, c0, c0, out_c, kActNone);
      v.neg_first = 3;
      v.res1_buf = in_buf; v.res1_lo_buf = lo(in_buf); v.res1_pitch = in_pitch; v.res1_nch = 3; v.beta1 = 1.f;
      if (last) {
        v.out_mode = c.out_fmt == 1 ? kOutNCHWF16 : kOutNCHWF32; v.out_buf = kBufExternalOut;
      } else {
        plain(v, t1out, round_up16(out_c));
      }
      P->add_conv(v);
    }
  };
  den_block("temp1.", in16, 16, 4, mid, false);
  den_block("temp2.", t1out, mid, mid, 3, true);

  // ---- owned frames + temporal halo: run every step only on the frames the owned outputs depend on.
  // Each shift conv reads its input at frames t-1, t, t+1, so a tensor is needed on the owned range widened by the
  // number of shift convs between it and the output: 16 for the first layers, 0 for the last -- about half the
  // halo work of a chunk with a 16-frame halo on each side (sharding.bsvd_chunks) falls away.
  const int own_lo = c.own_hi > c.own_lo ? c.own_lo : 0, own_hi = c.own_hi > c.own_lo ? c.own_hi : T;
  if (own_lo < 0 || own_hi > T) return "BSVD: owned frame range outside the clip";
  if (!c.bsvd_stream && (own_lo > 0 || own_hi < T)) {
    const int nb = static_cast<int>(P->bufs.size());
    std::vector<int> need(nb, -1), shifted(nb, 0);
    for (const Step& st : P->steps)
      if (st.kind == 1 && st.conv.out_buf >= 0 && st.conv.tshift) shifted[st.conv.out_buf] = 1;
    auto widen = [&](int buf, int r) { if (buf >= 0 && r > need[buf]) need[buf] = r; };
    for (int si = static_cast<int>(P->steps.size()) - 1; si >= 0; --si) {
      Step& st = P->steps[si];
      const int outb = st.kind == 1 ? st.conv.out_buf : st.prep.out_buf;
      const int R = outb >= 0 ? std::max(need[outb], 0) : 0;   // frames beyond the owned range this step must produce
      const int n0 = std::max(0, own_lo - R), n1 = std::min(T, own_hi + R);
      if (st.kind == 1) {
        ConvSpec& v = st.conv;
        P->flops -= v.flops();
        v.n0 = n0; v.n = n1 - n0; v.n_total = T;
        P->flops += v.flops();
        widen(v.in_buf, R + (shifted[v.in_buf] ? 1 : 0));
        widen(v.res1_buf, R);
        widen(v.res2_buf, R);
      } else {
        st.prep.n0 = n0; st.prep.n = n1 - n0;
      }
    }
    P->out_n = own_hi - own_lo;   // the result holds the owned frames only
  }
  return "";
}

}  // namespace ss4k
