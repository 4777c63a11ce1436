// iw3.depth_aa container (iw3/models/depth_aa.py:30-87, state_dict keys `proj_in.*`, `blocks.N.*`, `proj_out.*`): the learned
// anti-aliasing filter of Depth-Anything's output.  Included by model.cu.  The WindowScoreBias table is evaluated at pack time
// (pack_window_bias, rowflow_model.inl).
namespace nb200 {

struct AaBlockW {
    Lin qkv, proj, mlp0, mlp3;
    size_t bias = 0;   // fp32 [64][64]
    int shift = 0;
};
struct AaW {
    size_t win = 0, bin = 0;     // proj_in  fp32 [32][4], [32]
    size_t wout = 0, bout = 0;   // proj_out fp32 [4][32], [4]
    AaBlockW blk[3];
};

static std::shared_ptr<AaW> pack_depth_aa(Packer& pk) {
    auto r = std::make_shared<AaW>();
    if (const float* w = pk.get("proj_in.weight", 32 * 4)) r->win = pk.add_f32(std::vector<float>(w, w + 128));
    if (const float* b = pk.get("proj_in.bias", 32)) r->bin = pk.add_f32(std::vector<float>(b, b + 32));
    if (const float* w = pk.get("proj_out.weight", 4 * 32)) r->wout = pk.add_f32(std::vector<float>(w, w + 128));
    if (const float* b = pk.get("proj_out.bias", 4)) r->bout = pk.add_f32(std::vector<float>(b, b + 4));
    for (int i = 0; i < 3; ++i) {
        const std::string p = "blocks." + std::to_string(i) + ".";
        AaBlockW& b = r->blk[i];
        b.shift = i != 1;                                                   // depth_aa.py:38-42
        b.qkv = pack_linear(pk, p + "mha.mha.qkv_proj", 96, 32);
        b.proj = pack_linear(pk, p + "mha.mha.head_proj", 32, 32);
        b.mlp0 = pack_conv(pk, p + "conv_mlp.0", 32, 32, 1, 1);
        b.mlp3 = pack_conv(pk, p + "conv_mlp.3", 32, 32, 3, 3);
        b.bias = pack_window_bias(pk, p + "bias.", 8);
    }
    return r;
}

// DepthAA.forward (mode 0: eval clamp, mode 2: no clamp) / DepthAA.infer (mode 1: whole-tensor min/max normalisation, :46-55)
static int depth_aa_forward(nb200_model* m, cudaStream_t st, const float* x, int B, int H, int W, int mode, float* out) {
    const AaW& r = *m->aa;
    const int pad_w = 16 - W % 16, pad_h = 16 - H % 16;                     // always pads, also when already aligned (:61-62)
    const int pw1 = pad_w / 2, ph1 = pad_h / 2;
    const int Hh = (H + pad_h) / 2, Wh = (W + pad_w) / 2;
    const long long M = (long long)B * Hh * Wh;
    size_t bytes = 4096;
    auto need = [&](size_t elems) { bytes += ((elems * 2 + 255) & ~(size_t)255) + 256; };
    need((size_t)M * 32); need((size_t)M * 96); need((size_t)M * 32); need((size_t)M * 32); need((size_t)B * (Hh + 2) * (Wh + 2) * 32);
    if (m->ensure_ws(bytes)) return 1;
    Arena a{m->ws, 0, m->ws_bytes};
    float* mm = a.take<float>(64);
    __half* X = a.take<__half>((size_t)M * 32);
    __half* QKV = a.take<__half>((size_t)M * 96);
    __half* ATT = a.take<__half>((size_t)M * 32);
    __half* T = a.take<__half>((size_t)M * 32);
    __half* TP = a.take<__half>((size_t)B * (Hh + 2) * (Wh + 2) * 32);
    const float* mmp = nullptr;
    if (mode == 1) {
        if (aa_minmax(st, x, (long long)B * H * W, mm)) return 1;
        mmp = mm;
    }
    if (aa_prep(st, x, mmp, B, H, W, ph1, pw1, Hh, Wh, m->at<float>(r.win), m->at<float>(r.bin), X)) return 1;
    for (int i = 0; i < 3; ++i) {
        const AaBlockW& b = r.blk[i];
        // x = x + mha(x, attn_mask=bias)                                    depth_aa.py:24
        if (linear_flat(st, m, b.qkv, X, M, 32, QKV, 96, ACT_NONE)) return 1;
        if (aa_window_attention(st, QKV, m->at<float>(b.qkv.b), m->at<float>(b.bias), ATT, B, Hh, Wh, b.shift)) return 1;
        if (linear_flat(st, m, b.proj, ATT, M, 32, X, 32, ACT_NONE, X, 32)) return 1;
        // x = x + lrelu(conv3x3(reppad(gelu(conv1x1(x)))))                    :25
        if (linear_flat(st, m, b.mlp0, X, M, 32, T, 32, ACT_GELU)) return 1;
        if (aa_reppad(st, T, B, Hh, Wh, 32, TP)) return 1;
        ConvGemm g;
        g.A = TP; g.B = B; g.Hi = Hh + 2; g.Wi = Wh + 2; g.Ci = 32; g.Cin = 32; g.kind = CG_CONV3;
        g.Wt = m->at<__half>(b.mlp3.w); g.N = 32; g.bias = m->at<float>(b.mlp3.b); g.act = ACT_LRELU01; g.out = X; g.ldo = 32;
        g.res = X; g.ldr = 32; g.res_H = Hh; g.res_W = Wh;
        if (conv_gemm(st, g)) return 1;
    }
    return aa_out(st, X, x, mmp, B, H, W, ph1, pw1, Hh, Wh, m->at<float>(r.wout), m->at<float>(r.bout), mode == 0, out);
}

}  // namespace nb200
