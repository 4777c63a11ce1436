// Depth-Anything-V2 (DINOv2 ViT encoder + DPT head) container: weight packing from the upstream state_dict keys
// (`pretrained.*`, `depth_head.*`; what torch.hub "nagadomi/Depth-Anything_iw3" DepthAnything(encoder="v2_vits")
// loads, iw3/depth_anything_model.py:223-230) and the forward pass as a sequence of tcgen05 GEMMs and the kernels in
// depth_kernels.cu.  Included by model.cu.  Restated architecture + parity anchor: oracle/depth_anything.py.
//
// Pack-time algebra (exact in real arithmetic, fewer roundings than the reference's fp16 intermediates):
//   * LayerScale folded into the producing Linear:  gamma * (W x + b) = (gamma . W) x + gamma . b
//   * DPT "projects[i]" (1x1 conv) folded into the following ConvTranspose2d (k = s = 4 or 2): both are linear per pixel
// The residual stream is fp32 (as in the reference under autocast, where cat([cls fp32, tokens fp16]) promotes it);
// every Linear emits an fp16 delta that the fused add+LayerNorm kernel accumulates.
namespace nb200 {

struct DaBlockW {
    Lin qkv, proj, fc1, fc2;
    size_t n1w = 0, n1b = 0, n2w = 0, n2b = 0;
};
struct DaRefineW {
    Lin out_conv, c[2][2];   // resConfUnit{1,2}.conv{1,2}
};
struct DaW {
    int dim = 384, depth = 12, heads = 6, feat = 64, kpad = 640, pos_grid = 37;
    int c0pad = 64;    // channel stride of the first reassembled map (oc[0] rounded up to a multiple of 32)
    int oc[4] = {48, 96, 192, 384}, idx[4] = {2, 5, 8, 11};
    Lin patch, reasm[4], resize3, rn[4], oc1, oc2;
    DaRefineW ref[4];   // refinenet1..4
    size_t cls = 0, normw = 0, normb = 0, oc3w = 0;
    float oc3b = 0.f;
    std::vector<float> pos_host;   // learned table [1 + grid*grid][dim]
    std::vector<DaBlockW> blocks;
    // interpolated position table of the last token grid
    int pos_ph = 0, pos_pw = 0;
    float* pos_dev = nullptr;
    ~DaW() { if (pos_dev) cudaFree(pos_dev); }
};

static size_t pack_vec_f32(Packer& pk, const std::string& name, int n) {
    const float* v = pk.get(name, n);
    if (!v) return 0;
    return pk.add_f32(std::vector<float>(v, v + n));
}

// Linear [N][K] with a per-output scale folded in (LayerScale)
static Lin pack_linear_scaled(Packer& pk, const std::string& name, int N, int K, const std::string& gamma_name) {
    Lin l;
    l.N = N; l.K = K;
    const float* w = pk.get(name + ".weight", (int64_t)N * K);
    const float* b = pk.get(name + ".bias", N);
    const float* g = pk.get(gamma_name, N);
    if (!w || !b || !g) return l;
    std::vector<float> wv((size_t)N * K), bv(N);
    for (int n = 0; n < N; ++n) {
        for (int k = 0; k < K; ++k) wv[(size_t)n * K + k] = g[n] * w[(size_t)n * K + k];
        bv[n] = g[n] * b[n];
    }
    l.w = pk.add_f16(wv);
    l.b = pk.add_f32(bv);
    return l;
}

// projects[i] (Conv2d dim->c, k1) followed by ConvTranspose2d(c, c, k=r, s=r): one Linear dim -> r*r*c,
// rows n = (dy*r + dx)*c + co
static Lin pack_project_convT(Packer& pk, const std::string& proj, const std::string& convt, int dim, int c, int r) {
    Lin l;
    l.N = r * r * c; l.K = dim;
    const float* wp = pk.get(proj + ".weight", (int64_t)c * dim);
    const float* bp = pk.get(proj + ".bias", c);
    const float* wt = pk.get(convt + ".weight", (int64_t)c * c * r * r);   // [cin][cout][r][r]
    const float* bt = pk.get(convt + ".bias", c);
    if (!wp || !bp || !wt || !bt) return l;
    std::vector<float> wv((size_t)l.N * dim, 0.f), bv(l.N, 0.f);
    for (int g = 0; g < r * r; ++g)
        for (int co = 0; co < c; ++co) {
            float* row = &wv[((size_t)g * c + co) * dim];
            double bacc = bt[co];
            for (int m = 0; m < c; ++m) {
                const float t = wt[((size_t)m * c + co) * r * r + g];
                bacc += (double)t * bp[m];
                const float* src = wp + (size_t)m * dim;
                for (int k = 0; k < dim; ++k) row[k] += t * src[k];
            }
            bv[g * c + co] = (float)bacc;
        }
    l.w = pk.add_f16(wv);
    l.b = pk.add_f32(bv);
    return l;
}

// Conv2d without bias (scratch.layerN_rn)
static Lin pack_conv_nobias(Packer& pk, const std::string& name, int cout, int cin, int cin_pad) {
    Lin l;
    l.N = cout; l.K = 9 * cin_pad;
    const float* w = pk.get(name + ".weight", (int64_t)cout * cin * 9);
    if (!w) return l;
    std::vector<float> wv((size_t)cout * l.K, 0.f), bv(cout, 0.f);
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int k = 0; k < 9; ++k) wv[(size_t)co * l.K + (size_t)k * cin_pad + ci] = w[((size_t)co * cin + ci) * 9 + k];
    l.w = pk.add_f16(wv);
    l.b = pk.add_f32(bv);
    return l;
}

// encoder: 0 = ViT-S, 1 = ViT-B, 2 = ViT-L (Depth-Anything-V2 dpt.py model_configs / intermediate_layer_idx)
static std::shared_ptr<DaW> pack_depth_anything(Packer& pk, int encoder) {
    auto d = std::make_shared<DaW>();
    DaW& w = *d;
    if (encoder == 1) {
        w.dim = 768; w.depth = 12; w.heads = 12; w.feat = 128;
        const int oc[4] = {96, 192, 384, 768};
        for (int i = 0; i < 4; ++i) w.oc[i] = oc[i];
    } else if (encoder == 2) {
        w.dim = 1024; w.depth = 24; w.heads = 16; w.feat = 256;
        const int oc[4] = {256, 512, 1024, 1024}, idx[4] = {4, 11, 17, 23};
        for (int i = 0; i < 4; ++i) { w.oc[i] = oc[i]; w.idx[i] = idx[i]; }
    }
    w.c0pad = (w.oc[0] + 31) / 32 * 32;
    const int dim = w.dim;
    {   // patch embedding: Conv2d(3, dim, 14, 14) as a Linear over im2col rows, K = 588 zero-padded to 640
        Lin l;
        l.N = dim; l.K = w.kpad;
        const float* pw = pk.get("pretrained.patch_embed.proj.weight", (int64_t)dim * 588);
        const float* pb = pk.get("pretrained.patch_embed.proj.bias", dim);
        if (pw && pb) {
            std::vector<float> wv((size_t)dim * w.kpad, 0.f);
            for (int n = 0; n < dim; ++n) memcpy(&wv[(size_t)n * w.kpad], pw + (size_t)n * 588, 588 * 4);
            l.w = pk.add_f16(wv);
            l.b = pk.add_f32(std::vector<float>(pb, pb + dim));
        }
        w.patch = l;
    }
    w.cls = pack_vec_f32(pk, "pretrained.cls_token", dim);
    pk.mark("pretrained.mask_token");
    {
        auto it = pk.src.find("pretrained.pos_embed");
        if (it == pk.src.end()) { if (pk.err.empty()) pk.err = "missing key in state_dict: pretrained.pos_embed"; }
        else {
            const int64_t rows = it->second.numel / dim;
            const int g = (int)std::lround(std::sqrt((double)(rows - 1)));
            if (rows < 2 || (int64_t)g * g + 1 != rows || it->second.numel % dim) { if (pk.err.empty()) pk.err = "pretrained.pos_embed must be [1, 1+g*g, dim]"; }
            else {
                w.pos_grid = g;
                w.pos_host.assign(it->second.data, it->second.data + it->second.numel);
                it->second.used = true;
            }
        }
    }
    for (int i = 0; i < w.depth; ++i) {
        const std::string p = "pretrained.blocks." + std::to_string(i) + ".";
        DaBlockW b;
        b.n1w = pack_vec_f32(pk, p + "norm1.weight", dim);
        b.n1b = pack_vec_f32(pk, p + "norm1.bias", dim);
        b.qkv = pack_linear(pk, p + "attn.qkv", 3 * dim, dim);
        b.proj = pack_linear_scaled(pk, p + "attn.proj", dim, dim, p + "ls1.gamma");
        b.n2w = pack_vec_f32(pk, p + "norm2.weight", dim);
        b.n2b = pack_vec_f32(pk, p + "norm2.bias", dim);
        b.fc1 = pack_linear(pk, p + "mlp.fc1", 4 * dim, dim);
        b.fc2 = pack_linear_scaled(pk, p + "mlp.fc2", dim, 4 * dim, p + "ls2.gamma");
        w.blocks.push_back(b);
    }
    w.normw = pack_vec_f32(pk, "pretrained.norm.weight", dim);
    w.normb = pack_vec_f32(pk, "pretrained.norm.bias", dim);
    const std::string h = "depth_head.";
    w.reasm[0] = pack_project_convT(pk, h + "projects.0", h + "resize_layers.0", dim, w.oc[0], 4);
    w.reasm[1] = pack_project_convT(pk, h + "projects.1", h + "resize_layers.1", dim, w.oc[1], 2);
    w.reasm[2] = pack_conv(pk, h + "projects.2", w.oc[2], dim, 1, 1);
    w.reasm[3] = pack_conv(pk, h + "projects.3", w.oc[3], dim, 1, 1);
    w.resize3 = pack_conv(pk, h + "resize_layers.3", w.oc[3], w.oc[3], 3, 3);
    for (int i = 0; i < 4; ++i)
        w.rn[i] = pack_conv_nobias(pk, h + "scratch.layer" + std::to_string(i + 1) + "_rn", w.feat, w.oc[i], i == 0 ? w.c0pad : w.oc[i]);
    for (int r = 0; r < 4; ++r) {
        const std::string p = h + "scratch.refinenet" + std::to_string(r + 1) + ".";
        w.ref[r].out_conv = pack_conv(pk, p + "out_conv", w.feat, w.feat, 1, 1);
        for (int u = 0; u < 2; ++u)
            for (int c = 0; c < 2; ++c) {
                const std::string cn = p + "resConfUnit" + std::to_string(u + 1) + ".conv" + std::to_string(c + 1);
                if (r == 3 && u == 0) {   // refinenet4 has no second input: its resConfUnit1 is never evaluated
                    pk.mark(cn + ".weight");
                    pk.mark(cn + ".bias");
                    continue;
                }
                w.ref[r].c[u][c] = pack_conv(pk, cn, w.feat, w.feat, 3, 3);
            }
    }
    w.oc1 = pack_conv(pk, h + "scratch.output_conv1", w.feat / 2, w.feat, 3, 3);
    w.oc2 = pack_conv(pk, h + "scratch.output_conv2.0", 32, w.feat / 2, 3, 3);
    w.oc3w = pack_vec_f32(pk, h + "scratch.output_conv2.2.weight", 32);
    if (const float* b3 = pk.get(h + "scratch.output_conv2.2.bias", 1)) w.oc3b = b3[0];
    return d;
}

// dinov2 interpolate_pos_encoding (interpolate_offset 0.1): bicubic (A = -0.75, align_corners=False) resample of the
// grid x grid table with scale factors (ph+0.1)/grid, (pw+0.1)/grid; ATen upsample_bicubic2d arithmetic in fp32.
static void interp_pos_table(const DaW& w, int ph, int pw, std::vector<float>& out) {
    const int dim = w.dim, g = w.pos_grid;
    out.assign((size_t)(1 + ph * pw) * dim, 0.f);
    memcpy(out.data(), w.pos_host.data(), (size_t)dim * 4);
    if (ph == g && pw == g) {
        memcpy(out.data() + dim, w.pos_host.data() + dim, (size_t)g * g * dim * 4);
        return;
    }
    const float sy = (float)(1.0 / ((double)(ph + 0.1) / g)), sx = (float)(1.0 / ((double)(pw + 0.1) / g));
    auto cc1 = [](float x) { const float A = -0.75f; return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };
    auto cc2 = [](float x) { const float A = -0.75f; return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };
    auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
    const float* tab = w.pos_host.data() + dim;
    for (int oy = 0; oy < ph; ++oy) {
        const float ry = sy * ((float)oy + 0.5f) - 0.5f;
        const int iy = (int)std::floor(ry);
        const float ty = ry - (float)iy;
        const float cy[4] = {cc2(ty + 1.f), cc1(ty), cc1(1.f - ty), cc2((1.f - ty) + 1.f)};
        for (int ox = 0; ox < pw; ++ox) {
            const float rx = sx * ((float)ox + 0.5f) - 0.5f;
            const int ix = (int)std::floor(rx);
            const float tx = rx - (float)ix;
            const float cx[4] = {cc2(tx + 1.f), cc1(tx), cc1(1.f - tx), cc2((1.f - tx) + 1.f)};
            float* dst = &out[(size_t)(1 + oy * pw + ox) * dim];
            for (int c = 0; c < dim; ++c) {
                float acc = 0.f;
                for (int a = 0; a < 4; ++a) {
                    const float* row = tab + (size_t)clampi(iy - 1 + a, g - 1) * g * dim;
                    float h = 0.f;
                    for (int b = 0; b < 4; ++b) h += row[(size_t)clampi(ix - 1 + b, g - 1) * dim + c] * cx[b];
                    acc += h * cy[a];
                }
                dst[c] = acc;
            }
        }
    }
}

static int da_linear(cudaStream_t st, const nb200_model* m, const Lin& l, const __half* A, long long M, int lda, __half* out, int act) {
    return linear_flat(st, m, l, A, M, lda, out, l.N, act);
}

// NHWC conv3x3 pad 1 (or 1x1 when taps == 1) on [B][H][W][Cin] -> [B][H][W][N]
static int da_conv(cudaStream_t st, const nb200_model* m, const Lin& l, const __half* A, int B, int H, int W, int Ci, int Cin, __half* out,
                   int act, const __half* res, bool k3) {
    ConvGemm g;
    g.A = A; g.B = B; g.Hi = H; g.Wi = W; g.Ci = Ci; g.Cin = Cin; g.kind = k3 ? CG_CONV3 : CG_LINEAR_2D; g.pad = k3 ? 1 : 0;
    g.Wt = m->at<__half>(l.w); g.N = l.N; g.bias = m->at<float>(l.b); g.act = act; g.out = out; g.ldo = l.N;
    if (res) { g.res = res; g.ldr = l.N; g.res_H = H; g.res_W = W; }
    return conv_gemm(st, g);
}

// FeatureFusionBlock (dpt blocks.py): x0 (+ resConfUnit1(x1)) -> resConfUnit2 -> bilinear resize -> out_conv
static int da_fusion(cudaStream_t st, const nb200_model* m, const DaRefineW& r, int F, const __half* x0, const __half* x1, int B, int h, int w,
                     int oh, int ow, __half* t_relu, __half* t_c1, __half* t_sum, __half* t_u, __half* t_up, __half* out) {
    const long long n = (long long)B * h * w * F;
    const __half* cur = x0;
    if (x1) {
        // res = conv2(relu(conv1(relu(x1)))) + x1 ; output = x0 + res  ==  conv2(...) + (x0 + x1)
        if (da_relu_add(st, x1, x0, t_relu, t_sum, n)) return 1;
        if (da_conv(st, m, r.c[0][0], t_relu, B, h, w, F, F, t_c1, ACT_RELU, nullptr, true)) return 1;
        if (da_conv(st, m, r.c[0][1], t_c1, B, h, w, F, F, t_u, ACT_NONE, t_sum, true)) return 1;
        cur = t_u;
    }
    if (da_relu_add(st, cur, nullptr, t_relu, nullptr, n)) return 1;
    if (da_conv(st, m, r.c[1][0], t_relu, B, h, w, F, F, t_c1, ACT_RELU, nullptr, true)) return 1;
    __half* u2 = cur == t_u ? t_sum : t_u;   // t_sum is free again once the first unit has consumed it
    if (da_conv(st, m, r.c[1][1], t_c1, B, h, w, F, F, u2, ACT_NONE, cur, true)) return 1;
    if (da_upsample_bilinear(st, u2, B, h, w, F, t_up, oh, ow)) return 1;
    return da_conv(st, m, r.out_conv, t_up, B, oh, ow, F, F, out, ACT_NONE, nullptr, false);
}

static int depth_anything_forward(nb200_model* m, cudaStream_t st, const float* x, int B, int H, int W, float* depth) {
    DaW& w = *m->da;
    NB_CHECK(H % 14 == 0 && W % 14 == 0 && H >= 14 && W >= 14, "input height and width must be multiples of 14 (batch_preprocess)");
    const int dim = w.dim, ph = H / 14, pw = W / 14, P = ph * pw, N = P + 1, F = w.feat, F2 = w.feat / 2;
    const int c0 = w.oc[0], c0p = w.c0pad, c1 = w.oc[1], c2 = w.oc[2], c3 = w.oc[3];
    const long long M = (long long)B * N;
    if (w.pos_ph != ph || w.pos_pw != pw) {
        std::vector<float> tab;
        interp_pos_table(w, ph, pw, tab);
        if (w.pos_dev) cudaFree(w.pos_dev);
        w.pos_dev = nullptr;
        NB_CUDA(cudaMalloc((void**)&w.pos_dev, tab.size() * 4));
        NB_CUDA(cudaMemcpyAsync(w.pos_dev, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, st));
        NB_CUDA(cudaStreamSynchronize(st));   // `tab` is a host temporary (once per token-grid shape)
        w.pos_ph = ph; w.pos_pw = pw;
    }
    const int h1 = 4 * ph, w1 = 4 * pw, h2 = 2 * ph, w2 = 2 * pw, h3 = ph, w3 = pw, h4 = (ph + 1) / 2, w4 = (pw + 1) / 2;
    const int hp = 2 * h1, wp = 2 * w1;   // path_1 / output_conv1 resolution
    // ---- workspace
    size_t bytes = 4096;
    auto need = [&](size_t elems, size_t esz) { bytes += ((elems * esz + 255) & ~(size_t)255) + 256; };
    const size_t tcols = (size_t)16 * c0 > (size_t)dim ? (size_t)16 * c0 : (size_t)dim;
    need((size_t)B * P * w.kpad, 2); need((size_t)B * P * tcols, 2); need((size_t)M * dim, 4); need((size_t)M * dim, 2);
    need((size_t)M * 3 * dim, 2); need((size_t)M * dim, 2); need((size_t)M * 4 * dim, 2); need((size_t)M * dim, 2);
    for (int i = 0; i < 4; ++i) need((size_t)M * dim, 2);
    need((size_t)B * h1 * w1 * c0p, 2); need((size_t)B * h2 * w2 * c1, 2); need((size_t)B * h3 * w3 * c2, 2); need((size_t)B * h3 * w3 * c3, 2);
    need((size_t)B * h4 * w4 * 9 * c3, 2); need((size_t)B * h4 * w4 * c3, 2);
    need((size_t)B * h1 * w1 * F, 2); need((size_t)B * h2 * w2 * F, 2); need((size_t)B * h3 * w3 * F, 2); need((size_t)B * h4 * w4 * F, 2);
    for (int i = 0; i < 4; ++i) need((size_t)B * h1 * w1 * F, 2);           // relu / conv1 / sum / unit out (largest fusion resolution)
    need((size_t)B * hp * wp * F, 2);                                        // upsampled
    need((size_t)B * h3 * w3 * F, 2); need((size_t)B * h2 * w2 * F, 2); need((size_t)B * h1 * w1 * F, 2); need((size_t)B * hp * wp * F, 2);  // paths
    need((size_t)B * hp * wp * F2, 2); need((size_t)B * H * W * F2, 2); need((size_t)B * H * W * 32, 2);
    if (m->ensure_ws(bytes)) return 1;
    Arena a{m->ws, 0, m->ws_bytes};
    __half* Apatch = a.take<__half>((size_t)B * P * w.kpad);
    __half* T = a.take<__half>((size_t)B * P * tcols);   // patch GEMM output, later the reassemble-0 GEMM output
    float* X32 = a.take<float>((size_t)M * dim);
    __half* Hn = a.take<__half>((size_t)M * dim);
    __half* QKV = a.take<__half>((size_t)M * 3 * dim);
    __half* ATT = a.take<__half>((size_t)M * dim);
    __half* HID = a.take<__half>((size_t)M * 4 * dim);
    __half* D = a.take<__half>((size_t)M * dim);
    __half* FE[4];
    for (int i = 0; i < 4; ++i) FE[i] = a.take<__half>((size_t)M * dim);
    __half* L1 = a.take<__half>((size_t)B * h1 * w1 * c0p);
    __half* L2 = a.take<__half>((size_t)B * h2 * w2 * c1);
    __half* L3 = a.take<__half>((size_t)B * h3 * w3 * c2);
    __half* L4lin = a.take<__half>((size_t)B * h3 * w3 * c3);
    __half* L4col = a.take<__half>((size_t)B * h4 * w4 * 9 * c3);
    __half* L4 = a.take<__half>((size_t)B * h4 * w4 * c3);
    __half* R1 = a.take<__half>((size_t)B * h1 * w1 * F);
    __half* R2 = a.take<__half>((size_t)B * h2 * w2 * F);
    __half* R3 = a.take<__half>((size_t)B * h3 * w3 * F);
    __half* R4 = a.take<__half>((size_t)B * h4 * w4 * F);
    __half* t_relu = a.take<__half>((size_t)B * h1 * w1 * F);
    __half* t_c1 = a.take<__half>((size_t)B * h1 * w1 * F);
    __half* t_sum = a.take<__half>((size_t)B * h1 * w1 * F);
    __half* t_u = a.take<__half>((size_t)B * h1 * w1 * F);
    __half* t_up = a.take<__half>((size_t)B * hp * wp * F);
    __half* P4 = a.take<__half>((size_t)B * h3 * w3 * F);
    __half* P3 = a.take<__half>((size_t)B * h2 * w2 * F);
    __half* P2 = a.take<__half>((size_t)B * h1 * w1 * F);
    __half* P1 = a.take<__half>((size_t)B * hp * wp * F);
    __half* O1 = a.take<__half>((size_t)B * hp * wp * F2);
    __half* O1u = a.take<__half>((size_t)B * H * W * F2);
    __half* O2 = a.take<__half>((size_t)B * H * W * 32);

    // ---- encoder (dinov2 vision_transformer.py prepare_tokens_with_masks + blocks)
    if (da_patch_im2col(st, x, B, H, W, Apatch, w.kpad)) return 1;
    if (da_linear(st, m, w.patch, Apatch, (long long)B * P, w.kpad, T, ACT_NONE)) return 1;
    if (da_assemble_tokens(st, T, m->at<float>(w.cls), w.pos_dev, X32, B, P, dim)) return 1;
    const __half* pending = nullptr;
    int nf = 0;
    for (int i = 0; i < w.depth; ++i) {
        const DaBlockW& b = w.blocks[i];
        if (da_add_layernorm(st, X32, pending, m->at<float>(b.n1w), m->at<float>(b.n1b), Hn, M, dim)) return 1;
        if (da_linear(st, m, b.qkv, Hn, M, dim, QKV, ACT_NONE)) return 1;
        if (da_attention(st, QKV, ATT, B, N, w.heads)) return 1;
        if (da_linear(st, m, b.proj, ATT, M, dim, D, ACT_NONE)) return 1;                 // ls1 folded
        if (da_add_layernorm(st, X32, D, m->at<float>(b.n2w), m->at<float>(b.n2b), Hn, M, dim)) return 1;
        if (da_linear(st, m, b.fc1, Hn, M, dim, HID, ACT_GELU)) return 1;
        if (da_linear(st, m, b.fc2, HID, M, 4 * dim, D, ACT_NONE)) return 1;              // ls2 folded
        pending = D;
        if (nf < 4 && i == w.idx[nf]) {
            // get_intermediate_layers(..., norm=True): the final LayerNorm applied to this block's output
            if (da_add_layernorm(st, X32, D, m->at<float>(w.normw), m->at<float>(w.normb), FE[nf], M, dim)) return 1;
            pending = nullptr;
            ++nf;
        }
    }
    // ---- DPT head (dpt.py DPTHead.forward); the class token row of every image is skipped by the A view
    auto reassemble = [&](int i, __half* out, int out_mode, int cout) {
        ConvGemm g;
        g.A = FE[i] + dim; g.B = B; g.Hi = ph; g.Wi = pw; g.Ci = dim; g.Cin = dim; g.kind = CG_LINEAR_2D;
        g.a_row_stride = (long long)pw * dim; g.a_img_stride = (long long)N * dim;
        g.Wt = m->at<__half>(w.reasm[i].w); g.N = w.reasm[i].N; g.bias = m->at<float>(w.reasm[i].b); g.act = ACT_NONE;
        g.out = out; g.ldo = out_mode == OUT_PIXSHUF2 ? cout : w.reasm[i].N; g.out_mode = out_mode; g.cout = cout;
        return conv_gemm(st, g);
    };
    if (reassemble(0, T, 0, 0)) return 1;                                   // [B*P][16*48]
    if (da_depth_to_space4(st, T, B, ph, pw, c0, L1, c0p)) return 1;        // [B][4ph][4pw][c0 -> c0p]
    if (reassemble(1, L2, OUT_PIXSHUF2, c1)) return 1;                      // [B][2ph][2pw][c1]
    if (reassemble(2, L3, 0, 0)) return 1;
    if (reassemble(3, L4lin, 0, 0)) return 1;
    if (da_im2col_s2(st, L4lin, B, h3, w3, c3, L4col)) return 1;
    if (da_linear(st, m, w.resize3, L4col, (long long)B * h4 * w4, 9 * c3, L4, ACT_NONE)) return 1;
    if (da_conv(st, m, w.rn[0], L1, B, h1, w1, c0p, c0p, R1, ACT_NONE, nullptr, true)) return 1;
    if (da_conv(st, m, w.rn[1], L2, B, h2, w2, c1, c1, R2, ACT_NONE, nullptr, true)) return 1;
    if (da_conv(st, m, w.rn[2], L3, B, h3, w3, c2, c2, R3, ACT_NONE, nullptr, true)) return 1;
    if (da_conv(st, m, w.rn[3], L4, B, h4, w4, c3, c3, R4, ACT_NONE, nullptr, true)) return 1;
    if (da_fusion(st, m, w.ref[3], F, R4, nullptr, B, h4, w4, h3, w3, t_relu, t_c1, t_sum, t_u, t_up, P4)) return 1;
    if (da_fusion(st, m, w.ref[2], F, P4, R3, B, h3, w3, h2, w2, t_relu, t_c1, t_sum, t_u, t_up, P3)) return 1;
    if (da_fusion(st, m, w.ref[1], F, P3, R2, B, h2, w2, h1, w1, t_relu, t_c1, t_sum, t_u, t_up, P2)) return 1;
    if (da_fusion(st, m, w.ref[0], F, P2, R1, B, h1, w1, hp, wp, t_relu, t_c1, t_sum, t_u, t_up, P1)) return 1;
    if (da_conv(st, m, w.oc1, P1, B, hp, wp, F, F, O1, ACT_NONE, nullptr, true)) return 1;
    if (da_upsample_bilinear(st, O1, B, hp, wp, F2, O1u, H, W)) return 1;
    if (da_conv(st, m, w.oc2, O1u, B, H, W, F2, F2, O2, ACT_RELU, nullptr, true)) return 1;
    return da_head_final(st, O2, (long long)B * H * W, 32, m->at<float>(w.oc3w), w.oc3b, depth);
}

}  // namespace nb200
