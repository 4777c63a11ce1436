// sbs.mlbw container (iw3/models/mlbw.py:36-127; MLBW(num_layers = 2 | 4, base_dim = 32, small, hole_mask = False); state_dict keys
// `lv1_in.1.*`, `lv2.N.*`, `lv1_out.1.*`).  Included by model.cu.  num_layers is read off lv1_out.1.bias (2 L outputs), `small`
// (two blocks, shifted along x only, :53-57) off the absence of lv2.2.
namespace nb200 {

struct MlBlockW {
    Lin qkv, proj, mlp0, mlp3;
    size_t bias = 0;   // fp32 [16][16]
    int pad_y = 0, pad_x = 0;
};
struct MlW {
    int L = 2, C = 64, C1 = 8, nblk = 4;
    size_t win = 0, bin = 0, wout = 0, bout = 0;   // fp32 conv (1, 9) weights as stored
    MlBlockW blk[4];
};

static std::shared_ptr<MlW> pack_mlbw(Packer& pk) {
    auto r = std::make_shared<MlW>();
    auto it = pk.src.find("lv1_out.1.bias");
    if (it == pk.src.end() || (it->second.numel != 4 && it->second.numel != 8)) {
        pk.err = "sbs.mlbw: lv1_out.1.bias must have 2 * num_layers elements (num_layers 2 or 4; hole_mask models are not supported)";
        return r;
    }
    r->L = (int)it->second.numel / 2;
    r->C = 32 * r->L;
    r->C1 = r->C / 8;
    const bool small = pk.src.find("lv2.2.conv_mlp.0.bias") == pk.src.end();
    r->nblk = small ? 2 : 4;
    const int C = r->C, C1 = r->C1, L = r->L;
    if (const float* w = pk.get("lv1_in.1.weight", (int64_t)C1 * 27)) r->win = pk.add_f32(std::vector<float>(w, w + C1 * 27));
    if (const float* b = pk.get("lv1_in.1.bias", C1)) r->bin = pk.add_f32(std::vector<float>(b, b + C1));
    if (const float* w = pk.get("lv1_out.1.weight", (int64_t)2 * L * C1 * 9)) r->wout = pk.add_f32(std::vector<float>(w, w + 2 * L * C1 * 9));
    if (const float* b = pk.get("lv1_out.1.bias", 2 * L)) r->bout = pk.add_f32(std::vector<float>(b, b + 2 * L));
    for (int i = 0; i < r->nblk; ++i) {
        const std::string p = "lv2." + std::to_string(i) + ".";
        MlBlockW& b = r->blk[i];
        const bool shifted = i % 2 == 0;                                    // :53-64
        b.pad_x = shifted ? 2 : 0;
        b.pad_y = shifted && !small ? 2 : 0;
        b.qkv = pack_linear(pk, p + "mha.mha.qkv_proj", 3 * C, C);
        b.proj = pack_linear(pk, p + "mha.mha.head_proj", C, C);
        b.mlp0 = pack_conv(pk, p + "conv_mlp.0", C, C, 1, 1);
        b.mlp3 = pack_conv(pk, p + "conv_mlp.3", C, C, 3, 3);
        b.bias = pack_window_bias(pk, p + "bias.", 4);
    }
    return r;
}

// MLBW._forward in eval mode (:96-127): x fp32 [B][3][h][w] -> delta [B][L][h][w], layer_weight [B][L][h][w] (softmax over L)
static int mlbw_forward(nb200_model* m, cudaStream_t st, const float* x, int B, int h, int w, float* delta, float* lw) {
    const MlW& r = *m->ml;
    const int C = r.C;
    const int pad_w = 32 - w % 32, pad_h = 4 - h % 4;                       // _calc_pad :72-88 (always pads)
    const int pw1 = pad_w / 2, ph1 = pad_h / 2;
    const int Hp = h + pad_h, Wt = (w + pad_w) / 8;
    const long long M = (long long)B * Hp * Wt;
    size_t bytes = 4096;
    auto need = [&](size_t elems) { bytes += ((elems * 2 + 255) & ~(size_t)255) + 256; };
    need((size_t)M * C); need((size_t)M * C); need((size_t)M * 3 * C); need((size_t)M * C); need((size_t)M * C);
    need((size_t)B * (Hp + 2) * (Wt + 2) * C);
    if (m->ensure_ws(bytes)) return 1;
    Arena a{m->ws, 0, m->ws_bytes};
    __half* T0 = a.take<__half>((size_t)M * C);
    __half* X = a.take<__half>((size_t)M * C);
    __half* QKV = a.take<__half>((size_t)M * 3 * C);
    __half* ATT = a.take<__half>((size_t)M * C);
    __half* T = a.take<__half>((size_t)M * C);
    __half* TP = a.take<__half>((size_t)B * (Hp + 2) * (Wt + 2) * C);
    if (mlbw_prep(st, x, B, h, w, ph1, pw1, Hp, Wt, r.C1, m->at<float>(r.win), m->at<float>(r.bin), T0)) return 1;
    NB_CUDA(cudaMemcpyAsync(X, T0, (size_t)M * C * 2, cudaMemcpyDeviceToDevice, st));
    for (int i = 0; i < r.nblk; ++i) {
        const MlBlockW& b = r.blk[i];
        // x = x + mha(x, attn_mask=bias)                                    mlbw.py:31
        if (linear_flat(st, m, b.qkv, X, M, C, QKV, 3 * C, ACT_NONE)) return 1;
        if (mlbw_window_attention(st, QKV, m->at<float>(b.qkv.b), m->at<float>(b.bias), ATT, B, Hp, Wt, r.L, b.pad_y, b.pad_x)) return 1;
        if (linear_flat(st, m, b.proj, ATT, M, C, X, C, ACT_NONE, X, C)) return 1;
        // x = x + conv3x3(reppad(gelu(conv1x1(x))))                          :32 (no activation after the 3x3)
        if (linear_flat(st, m, b.mlp0, X, M, C, T, C, ACT_GELU)) return 1;
        if (aa_reppad(st, T, B, Hp, Wt, C, TP)) return 1;
        ConvGemm g;
        g.A = TP; g.B = B; g.Hi = Hp + 2; g.Wi = Wt + 2; g.Ci = C; g.Cin = C; g.kind = CG_CONV3;
        g.Wt = m->at<__half>(b.mlp3.w); g.N = C; g.bias = m->at<float>(b.mlp3.b); g.act = ACT_NONE; g.out = X; g.ldo = C;
        g.res = X; g.ldr = C; g.res_H = Hp; g.res_W = Wt;
        if (conv_gemm(st, g)) return 1;
    }
    return mlbw_out(st, X, T0, B, h, w, ph1, pw1, Hp, Wt, r.C1, r.L, m->at<float>(r.wout), m->at<float>(r.bout), delta, lw);
}

}  // namespace nb200
