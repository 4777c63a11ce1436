// sbs.row_flow_v3 container (iw3/models/row_flow_v3.py:32-68, state_dict keys `blocks.*`, `last_layer.1.*`): the learned
// delta network of iw3's default stereo method.  Included by model.cu.  The WindowScoreBias MLP (nunif/modules/attention.py
// :375-420) only depends on weights, so its (N x N) table is evaluated once at pack time.
namespace nb200 {

struct RfBlockW {
    Lin qkv, proj, mlp0, mlp3;
    size_t bias = 0;   // fp32 [N][N]
    int ws = 4;
};
struct RfW {
    Lin c0;            // Conv2d(24, 64, 1) as a Linear over K = 32 (zero padded)
    RfBlockW blk[2];
    size_t lastw = 0;  // fp32 [8][3][3]
    float lastb = 0.f;
};

static size_t pack_window_bias(Packer& pk, const std::string& p, int ws) {
    const int N = ws * ws, U = (2 * ws - 1) * (2 * ws - 1), hid = (int)std::sqrt((double)N) * 2;
    const float* idx = pk.get(p + "index", (int64_t)N * N);          // int64 buffer, delivered as float32 by the loader
    const float* dl = pk.get(p + "delta", (int64_t)U * 2);
    const float* w0 = pk.get(p + "to_bias.0.weight", (int64_t)hid * 2);
    const float* b0 = pk.get(p + "to_bias.0.bias", hid);
    const float* w1 = pk.get(p + "to_bias.2.weight", hid);
    const float* b1 = pk.get(p + "to_bias.2.bias", 1);
    if (!idx || !dl || !w0 || !b0 || !w1 || !b1) return 0;
    std::vector<float> tab(U), out((size_t)N * N);
    for (int u = 0; u < U; ++u) {
        double acc = b1[0];
        for (int j = 0; j < hid; ++j) {
            const double z = (double)w0[j * 2] * dl[u * 2] + (double)w0[j * 2 + 1] * dl[u * 2 + 1] + b0[j];
            acc += (double)w1[j] * (0.5 * z * (1.0 + std::erf(z * 0.7071067811865476)));        // nn.GELU (erf)
        }
        tab[u] = (float)acc;
    }
    for (int i = 0; i < N * N; ++i) {
        const int u = (int)std::lround(idx[i]);
        if (u < 0 || u >= U) { if (pk.err.empty()) pk.err = "bad window bias index in " + p; return 0; }
        out[i] = tab[u];
    }
    return pk.add_f32(out);
}

static std::shared_ptr<RfW> pack_row_flow(Packer& pk) {
    auto r = std::make_shared<RfW>();
    {
        Lin l;
        l.N = 64; l.K = 32;
        const float* w = pk.get("blocks.0.weight", 64 * 24);
        const float* b = pk.get("blocks.0.bias", 64);
        if (w && b) {
            std::vector<float> wv(64 * 32, 0.f);
            for (int n = 0; n < 64; ++n) memcpy(&wv[(size_t)n * 32], w + (size_t)n * 24, 24 * 4);
            l.w = pk.add_f16(wv);
            l.b = pk.add_f32(std::vector<float>(b, b + 64));
        }
        r->c0 = l;
    }
    for (int i = 0; i < 2; ++i) {
        const std::string p = "blocks." + std::to_string(i + 1) + ".";
        RfBlockW& b = r->blk[i];
        b.ws = i == 0 ? 4 : 3;
        b.qkv = pack_linear(pk, p + "mha.mha.qkv_proj", 192, 64);
        b.proj = pack_linear(pk, p + "mha.mha.head_proj", 64, 64);
        b.mlp0 = pack_conv(pk, p + "conv_mlp.0", 64, 64, 1, 1);
        b.mlp3 = pack_conv(pk, p + "conv_mlp.3", 64, 64, 3, 3);
        b.bias = pack_window_bias(pk, p + "bias.", b.ws);
    }
    if (const float* w = pk.get("last_layer.1.weight", 72)) r->lastw = pk.add_f32(std::vector<float>(w, w + 72));
    if (const float* b = pk.get("last_layer.1.bias", 1)) r->lastb = b[0];
    pk.mark("delta_scale");
    return r;
}

// RowFlowV3._forward (delta_output mode): x fp32 [B][3][h][w] (depth, divergence feature, convergence feature) -> delta [B][1][h][w]
static int row_flow_forward(nb200_model* m, cudaStream_t st, const float* x, int B, int h, int w, float* delta) {
    const RfW& r = *m->rf;
    const int pad1 = 96 - w % 96, pad2 = 12 - h % 12;          // row_flow_v3.py:59-60 (always pads, also when already aligned)
    const int Hp = h + pad2, Wp = w + pad1, Wt = Wp / 8;
    const long long M = (long long)B * Hp * Wt;
    size_t bytes = 4096;
    auto need = [&](size_t elems) { bytes += ((elems * 2 + 255) & ~(size_t)255) + 256; };
    need((size_t)M * 32); need((size_t)M * 64); need((size_t)M * 192); need((size_t)M * 64); need((size_t)M * 64);
    need((size_t)B * (Hp + 2) * (Wt + 2) * 64);
    if (m->ensure_ws(bytes)) return 1;
    Arena a{m->ws, 0, m->ws_bytes};
    __half* A0 = a.take<__half>((size_t)M * 32);
    __half* X = a.take<__half>((size_t)M * 64);
    __half* QKV = a.take<__half>((size_t)M * 192);
    __half* ATT = a.take<__half>((size_t)M * 64);
    __half* T = a.take<__half>((size_t)M * 64);
    __half* TP = a.take<__half>((size_t)B * (Hp + 2) * (Wt + 2) * 64);
    if (rf_prep(st, x, B, h, w, Hp, Wt, A0)) return 1;
    if (linear_flat(st, m, r.c0, A0, M, 32, X, 64, ACT_NONE)) return 1;
    for (int i = 0; i < 2; ++i) {
        const RfBlockW& b = r.blk[i];
        // x = x + mha(x, attn_mask=bias)                                    row_flow_v3.py:27
        if (linear_flat(st, m, b.qkv, X, M, 64, QKV, 192, ACT_NONE)) return 1;
        if (rf_window_attention(st, QKV, m->at<float>(b.bias), ATT, B, Hp, Wt, b.ws)) return 1;
        if (linear_flat(st, m, b.proj, ATT, M, 64, X, 64, ACT_NONE, X, 64)) return 1;
        // x = x + lrelu(conv3x3(reppad(gelu(conv1x1(x)))))                    :28
        if (linear_flat(st, m, b.mlp0, X, M, 64, T, 64, ACT_GELU)) return 1;
        if (rf_reppad(st, T, B, Hp, Wt, TP)) return 1;
        ConvGemm g;
        g.A = TP; g.B = B; g.Hi = Hp + 2; g.Wi = Wt + 2; g.Ci = 64; g.Cin = 64; g.kind = CG_CONV3;
        g.Wt = m->at<__half>(b.mlp3.w); g.N = 64; g.bias = m->at<float>(b.mlp3.b); g.act = ACT_LRELU01; g.out = X; g.ldo = 64;
        g.res = X; g.ldr = 64; g.res_H = Hp; g.res_W = Wt;
        if (conv_gemm(st, g)) return 1;
    }
    return rf_last_conv(st, X, B, Hp, Wt, h, w, m->at<float>(r.lastw), r.lastb, delta);
}

}  // namespace nb200
