// ZoeD_N (BEiT-L/16 encoder + MiDaS DPT head + ZoeDepth metric bins head) container: weight packing from the upstream
// checkpoint keys of ZoeD_M12_N.pt (`core.core.pretrained.*`, `core.core.scratch.*`, `conv2`, `seed_bin_regressor`,
// `seed_projector`, `projectors`, `attractors`, `conditional_log_binomial`; what torch.hub "nagadomi/ZoeDepth_iw3" ZoeD_N
// loads, iw3/zoedepth_model.py:151-157) and the forward pass as a sequence of tcgen05 GEMMs, the flash attention with the
// relative-position bias, and the kernels in depth_kernels.cu / zoe_kernels.cu.  Included by model.cu after depth_model.inl
// (shares its packing helpers and da_fusion).  Restated architecture + parity anchor: oracle/zoedepth.py.
//
// The configuration (embedding dim, depth, DPT widths, training grid of the relative-position table) is read off the
// tensor sizes, so the reduced test configuration (synth.ZOED_MINI) runs the same code.
// Pack-time algebra as in depth_model.inl: gamma_1 / gamma_2 folded into attn.proj / mlp.fc2, act_postprocess{1,2}.3 (1x1 conv)
// folded into the following ConvTranspose2d; q_bias | 0 | v_bias become one qkv bias vector (BEiT has no key bias).
namespace nb200 {

struct ZoeBlockW {
    Lin qkv, proj, fc1, fc2;
    size_t n1w = 0, n1b = 0, n2w = 0, n2b = 0;
};
struct ZoeW {
    int dim = 1024, depth = 24, heads = 16, feat = 256, old_grid = 24;
    int oc[4] = {256, 512, 1024, 1024}, hooks[4] = {5, 11, 17, 23};
    Lin patch, readout[4], reasm[4], resize3, rn[4], oc1, oc2;
    DaRefineW ref[4];
    size_t cls = 0, oc3w = 0;
    float oc3b = 0.f;
    Lin conv2, seed1, seed2, sproj1, sproj2, proj1[4], proj2[4], att1[4], att2[4], clb1;
    size_t clb2w = 0, clb2b = 0;
    int n_att[4] = {16, 8, 4, 1};
    std::vector<ZoeBlockW> blocks;
    std::vector<std::vector<float>> tables;   // per block: learned relative_position_bias_table [(2g-1)^2 + 3][heads]
    // expanded bias of the last token grid: [depth][heads][N][ldb] fp32, log2(e) folded in
    int bias_ph = 0, bias_pw = 0, ldb = 0;
    float* bias_dev = nullptr;
    ~ZoeW() { if (bias_dev) cudaFree(bias_dev); }
};

static int64_t zoe_numel(Packer& pk, const std::string& name) {
    auto it = pk.src.find(name);
    if (it == pk.src.end()) {
        if (pk.err.empty()) pk.err = "missing key in state_dict: " + name;
        return 0;
    }
    return it->second.numel;
}

// attn.qkv.weight [3 dim][dim] + (q_bias | 0 | v_bias)
static Lin pack_beit_qkv(Packer& pk, const std::string& p, int dim) {
    Lin l;
    l.N = 3 * dim; l.K = dim;
    const float* w = pk.get(p + "attn.qkv.weight", (int64_t)3 * dim * dim);
    const float* qb = pk.get(p + "attn.q_bias", dim);
    const float* vb = pk.get(p + "attn.v_bias", dim);
    if (!w || !qb || !vb) return l;
    std::vector<float> bv((size_t)3 * dim, 0.f);
    memcpy(bv.data(), qb, (size_t)dim * 4);
    memcpy(bv.data() + 2 * (size_t)dim, vb, (size_t)dim * 4);
    l.w = pk.add_f16(std::vector<float>(w, w + (size_t)3 * dim * dim));
    l.b = pk.add_f32(bv);
    return l;
}

static std::shared_ptr<ZoeW> pack_zoedepth(Packer& pk) {
    auto z = std::make_shared<ZoeW>();
    ZoeW& w = *z;
    const std::string bb = "core.core.pretrained.model.", pp = "core.core.pretrained.", sc = "core.core.scratch.";
    w.dim = (int)zoe_numel(pk, bb + "cls_token");
    if (!pk.err.empty()) return z;
    const int dim = w.dim;
    if (dim < 64 || dim % 64) { pk.err = "ZoeDepth: embedding dim must be a multiple of 64 (head dim 64)"; return z; }
    w.heads = dim / 64;
    w.depth = 0;
    while (pk.src.count(bb + "blocks." + std::to_string(w.depth) + ".norm1.weight")) ++w.depth;
    if (w.depth < 4 || w.depth % 4) { pk.err = "ZoeDepth: the encoder depth must be a positive multiple of 4"; return z; }
    for (int i = 0; i < 4; ++i) {
        w.hooks[i] = w.depth / 4 * (i + 1) - 1;     // BEiT-L: blocks 5, 11, 17, 23 (MiDaS dpt_depth.py hooks)
        w.oc[i] = (int)zoe_numel(pk, pp + "act_postprocess" + std::to_string(i + 1) + ".3.bias");
        if (pk.err.empty() && (w.oc[i] < 32 || w.oc[i] % 32)) pk.err = "ZoeDepth: reassemble widths must be multiples of 32";
    }
    if (!pk.err.empty()) return z;
    w.feat = (int)(zoe_numel(pk, sc + "layer1_rn.weight") / ((int64_t)w.oc[0] * 9));
    if (w.feat < 64 || w.feat % 64) { pk.err = "ZoeDepth: fusion width must be a multiple of 64"; return z; }
    {
        const int64_t rows = zoe_numel(pk, bb + "blocks.0.attn.relative_position_bias_table") / w.heads;
        const int s = (int)std::lround(std::sqrt((double)(rows - 3)));
        if (rows < 4 || (int64_t)s * s + 3 != rows || s % 2 == 0) { pk.err = "ZoeDepth: relative_position_bias_table must be [(2g-1)^2 + 3, heads]"; return z; }
        w.old_grid = (s + 1) / 2;
    }
    const int F = w.feat;
    {   // patch embedding: Conv2d(3, dim, 16, 16) as a Linear over im2col rows (K = 768)
        Lin l;
        l.N = dim; l.K = 768;
        const float* pw = pk.get(bb + "patch_embed.proj.weight", (int64_t)dim * 768);
        const float* pb = pk.get(bb + "patch_embed.proj.bias", dim);
        if (pw && pb) {
            l.w = pk.add_f16(std::vector<float>(pw, pw + (size_t)dim * 768));
            l.b = pk.add_f32(std::vector<float>(pb, pb + dim));
        }
        w.patch = l;
    }
    w.cls = pack_vec_f32(pk, bb + "cls_token", dim);
    const int64_t trows = (int64_t)(2 * w.old_grid - 1) * (2 * w.old_grid - 1) + 3;
    for (int i = 0; i < w.depth; ++i) {
        const std::string p = bb + "blocks." + std::to_string(i) + ".";
        ZoeBlockW b;
        b.n1w = pack_vec_f32(pk, p + "norm1.weight", dim);
        b.n1b = pack_vec_f32(pk, p + "norm1.bias", dim);
        b.qkv = pack_beit_qkv(pk, p, dim);
        b.proj = pack_linear_scaled(pk, p + "attn.proj", dim, dim, p + "gamma_1");
        b.n2w = pack_vec_f32(pk, p + "norm2.weight", dim);
        b.n2b = pack_vec_f32(pk, p + "norm2.bias", dim);
        b.fc1 = pack_linear(pk, p + "mlp.fc1", 4 * dim, dim);
        b.fc2 = pack_linear_scaled(pk, p + "mlp.fc2", dim, 4 * dim, p + "gamma_2");
        const float* t = pk.get(p + "attn.relative_position_bias_table", trows * w.heads);
        w.tables.emplace_back(t ? std::vector<float>(t, t + trows * w.heads) : std::vector<float>());
        pk.mark(p + "attn.relative_position_index");   // a buffer of the timm module; recomputed for the actual grid
        w.blocks.push_back(b);
    }
    for (const char* k : {"norm.weight", "norm.bias", "fc_norm.weight", "fc_norm.bias", "head.weight", "head.bias"}) pk.mark(bb + k);  // unused by DPT
    for (int i = 0; i < 4; ++i) {
        const std::string p = pp + "act_postprocess" + std::to_string(i + 1) + ".";
        w.readout[i] = pack_linear(pk, p + "0.project.0", dim, 2 * dim);
    }
    w.reasm[0] = pack_project_convT(pk, pp + "act_postprocess1.3", pp + "act_postprocess1.4", dim, w.oc[0], 4);
    w.reasm[1] = pack_project_convT(pk, pp + "act_postprocess2.3", pp + "act_postprocess2.4", dim, w.oc[1], 2);
    w.reasm[2] = pack_conv(pk, pp + "act_postprocess3.3", w.oc[2], dim, 1, 1);
    w.reasm[3] = pack_conv(pk, pp + "act_postprocess4.3", w.oc[3], dim, 1, 1);
    w.resize3 = pack_conv(pk, pp + "act_postprocess4.4", w.oc[3], w.oc[3], 3, 3);
    for (int i = 0; i < 4; ++i) w.rn[i] = pack_conv_nobias(pk, sc + "layer" + std::to_string(i + 1) + "_rn", F, w.oc[i], w.oc[i]);
    for (int r = 0; r < 4; ++r) {
        const std::string p = sc + "refinenet" + std::to_string(r + 1) + ".";
        w.ref[r].out_conv = pack_conv(pk, p + "out_conv", F, F, 1, 1);
        for (int u = 0; u < 2; ++u)
            for (int c = 0; c < 2; ++c) {
                const std::string cn = p + "resConfUnit" + std::to_string(u + 1) + ".conv" + std::to_string(c + 1);
                if (r == 3 && u == 0) {   // refinenet4 has no second input: its resConfUnit1 is never evaluated
                    pk.mark(cn + ".weight");
                    pk.mark(cn + ".bias");
                    continue;
                }
                w.ref[r].c[u][c] = pack_conv(pk, cn, F, F, 3, 3);
            }
    }
    w.oc1 = pack_conv(pk, sc + "output_conv.0", F / 2, F, 3, 3);
    w.oc2 = pack_conv(pk, sc + "output_conv.2", 32, F / 2, 3, 3);
    w.oc3w = pack_vec_f32(pk, sc + "output_conv.4.weight", 32);
    if (const float* b3 = pk.get(sc + "output_conv.4.bias", 1)) w.oc3b = b3[0];
    // metric bins head (zoedepth_v1.py): n_bins 64, bin_embedding_dim 128, attractors 16 / 8 / 4 / 1
    w.conv2 = pack_conv(pk, "conv2", F, F, 1, 1);
    w.seed1 = pack_conv(pk, "seed_bin_regressor._net.0", 256, F, 1, 1);
    w.seed2 = pack_conv(pk, "seed_bin_regressor._net.2", 64, 256, 1, 1);
    w.sproj1 = pack_conv(pk, "seed_projector._net.0", 128, F, 1, 1);
    w.sproj2 = pack_conv(pk, "seed_projector._net.2", 128, 128, 1, 1);
    for (int i = 0; i < 4; ++i) {
        const std::string s = std::to_string(i);
        w.proj1[i] = pack_conv(pk, "projectors." + s + "._net.0", 128, F, 1, 1);
        w.proj2[i] = pack_conv(pk, "projectors." + s + "._net.2", 128, 128, 1, 1);
        w.att1[i] = pack_conv(pk, "attractors." + s + "._net.0", 128, 128, 1, 1);
        w.att2[i] = pack_conv(pk, "attractors." + s + "._net.2", w.n_att[i], 128, 1, 1, 0, /*cout_pad=*/16);
    }
    // ConditionalLogBinomial mlp: (32 + 1 + 128 = 161 -> K padded to 192) -> 80 (N padded to 96) -> GELU -> 4
    // input channels are re-ordered [embedding 128 | activation 32 | relative depth | 31 zeros] so that zoe_clb_concat moves
    // whole 16-byte groups (the reference concatenates [activation | relative depth | embedding])
    {
        Lin l;
        l.N = 96; l.K = 192;
        const float* cw = pk.get("conditional_log_binomial.mlp.0.weight", (int64_t)80 * 161);
        const float* cb = pk.get("conditional_log_binomial.mlp.0.bias", 80);
        if (cw && cb) {
            std::vector<float> wv((size_t)96 * 192, 0.f), bv(96, 0.f);
            for (int n = 0; n < 80; ++n) {
                for (int j = 0; j < 128; ++j) wv[(size_t)n * 192 + j] = cw[(size_t)n * 161 + 33 + j];
                for (int j = 0; j < 33; ++j) wv[(size_t)n * 192 + 128 + j] = cw[(size_t)n * 161 + j];
                bv[n] = cb[n];
            }
            l.w = pk.add_f16(wv);
            l.b = pk.add_f32(bv);
        }
        w.clb1 = l;
    }
    if (const float* c2 = pk.get("conditional_log_binomial.mlp.2.weight", 4 * 80)) {
        std::vector<float> v(c2, c2 + 320);
        for (auto& x : v) x = __half2float(__float2half_rn(x));   // the reference runs this conv in fp16
        w.clb2w = pk.add_f32(v);
    }
    w.clb2b = pack_vec_f32(pk, "conditional_log_binomial.mlp.2.bias", 4);
    return z;
}

// MiDaS beit.py _get_rel_pos_bias: the (2g-1) x (2g-1) learned sub-table -> (2ph-1) x (2pw-1) by ATen's bilinear resample
// (align_corners=False, no antialias), the 3 class-token rows appended unchanged.  out: [(2ph-1)(2pw-1) + 3][heads]
static void zoe_resample_table(const std::vector<float>& tab, int g, int heads, int ph, int pw, float* out) {
    const int S = 2 * g - 1, nh = 2 * ph - 1, nw = 2 * pw - 1;
    const float sy = (float)S / (float)nh, sx = (float)S / (float)nw;
    for (int y = 0; y < nh; ++y) {
        float fy = sy * ((float)y + 0.5f) - 0.5f;
        if (fy < 0.f) fy = 0.f;
        const int y0 = (int)fy, y1 = y0 + (y0 < S - 1 ? 1 : 0);
        const float ly = fy - (float)y0, hy = 1.f - ly;
        for (int x = 0; x < nw; ++x) {
            float fx = sx * ((float)x + 0.5f) - 0.5f;
            if (fx < 0.f) fx = 0.f;
            const int x0 = (int)fx, x1 = x0 + (x0 < S - 1 ? 1 : 0);
            const float lx = fx - (float)x0, hx = 1.f - lx;
            const float* a = &tab[((size_t)y0 * S + x0) * heads];
            const float* b = &tab[((size_t)y0 * S + x1) * heads];
            const float* c = &tab[((size_t)y1 * S + x0) * heads];
            const float* d = &tab[((size_t)y1 * S + x1) * heads];
            float* o = out + ((size_t)y * nw + x) * heads;
            for (int h = 0; h < heads; ++h) o[h] = hy * (hx * a[h] + lx * b[h]) + ly * (hx * c[h] + lx * d[h]);
        }
    }
    memcpy(out + (size_t)nh * nw * heads, &tab[(size_t)S * S * heads], (size_t)3 * heads * 4);
}

static int zoe_prepare_bias(ZoeW& w, cudaStream_t st, int ph, int pw) {
    if (w.bias_dev && w.bias_ph == ph && w.bias_pw == pw) return 0;
    const int N = ph * pw + 1, ldb = (N + 63) / 64 * 64;
    const size_t rows = (size_t)(2 * ph - 1) * (2 * pw - 1) + 3, per_layer = (size_t)w.heads * N * ldb;
    std::vector<float> host(rows * w.heads * w.depth);
    for (int i = 0; i < w.depth; ++i) zoe_resample_table(w.tables[i], w.old_grid, w.heads, ph, pw, host.data() + (size_t)i * rows * w.heads);
    if (w.bias_dev) cudaFree(w.bias_dev);
    w.bias_dev = nullptr; w.bias_ph = w.bias_pw = 0;
    float* tdev = nullptr;
    NB_CUDA(cudaMalloc((void**)&w.bias_dev, per_layer * w.depth * 4));
    NB_CUDA(cudaMalloc((void**)&tdev, host.size() * 4));
    cudaError_t e = cudaMemsetAsync(w.bias_dev, 0, per_layer * w.depth * 4, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(tdev, host.data(), host.size() * 4, cudaMemcpyHostToDevice, st);
    int rc = 0;
    for (int i = 0; i < w.depth && e == cudaSuccess && !rc; ++i)
        rc = zoe_expand_rel_bias(st, tdev + (size_t)i * rows * w.heads, ph, pw, w.heads, w.bias_dev + (size_t)i * per_layer, ldb);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);   // `host` / `tdev` are temporaries (once per token-grid shape)
    cudaFree(tdev);
    if (e != cudaSuccess) return fail(std::string("zoe_prepare_bias: ") + cudaGetErrorString(e));
    if (rc) return 1;
    w.bias_ph = ph; w.bias_pw = pw; w.ldb = ldb;
    return 0;
}

static int zoedepth_forward(nb200_model* m, cudaStream_t st, const float* x, int B, int H, int W, float* depth) {
    ZoeW& w = *m->zoe;
    NB_CHECK(H % 32 == 0 && W % 32 == 0 && H >= 32 && W >= 32, "input height and width must be multiples of 32 (batch_preprocess)");
    const int dim = w.dim, ph = H / 16, pw = W / 16, P = ph * pw, N = P + 1, F = w.feat, F2 = w.feat / 2;
    const int c0 = w.oc[0], c1 = w.oc[1], c2 = w.oc[2], c3 = w.oc[3];
    const long long M = (long long)B * N;
    if (zoe_prepare_bias(w, st, ph, pw)) return 1;
    const size_t bias_layer = (size_t)w.heads * N * w.ldb;
    const int h1 = 4 * ph, w1 = 4 * pw, h2 = 2 * ph, w2 = 2 * pw, h3 = ph, w3 = pw, h4 = ph / 2, w4 = pw / 2;
    const int hp = 2 * h1, wp = 2 * w1;   // path_1 resolution = H/2 x W/2
    const size_t npix = (size_t)B * H * W;
    // ---- workspace
    size_t bytes = 4096;
    auto need = [&](size_t elems, size_t esz) { bytes += ((elems * esz + 255) & ~(size_t)255) + 256; };
    const size_t tcols = (size_t)16 * c0 > (size_t)2 * dim ? (size_t)16 * c0 : (size_t)2 * dim;
    need((size_t)B * P * 768, 2); need((size_t)B * P * tcols, 2); need((size_t)M * dim, 4); need((size_t)M * dim, 2);
    need((size_t)M * 3 * dim, 2); need((size_t)M * dim, 2); need((size_t)M * 4 * dim, 2); need((size_t)M * dim, 2);
    for (int i = 0; i < 4; ++i) need((size_t)M * dim, 2);
    need((size_t)B * P * dim, 2);
    need((size_t)B * h1 * w1 * c0, 2); need((size_t)B * h2 * w2 * c1, 2); need((size_t)B * h3 * w3 * c2, 2); need((size_t)B * h3 * w3 * c3, 2);
    need((size_t)B * h4 * w4 * 9 * c3, 2); need((size_t)B * h4 * w4 * c3, 2);
    need((size_t)B * h1 * w1 * F, 2); need((size_t)B * h2 * w2 * F, 2); need((size_t)B * h3 * w3 * F, 2); need((size_t)B * h4 * w4 * F, 2);
    for (int i = 0; i < 4; ++i) need((size_t)B * h1 * w1 * F, 2);
    need((size_t)B * hp * wp * F, 2);
    need((size_t)B * h3 * w3 * F, 2); need((size_t)B * h2 * w2 * F, 2); need((size_t)B * h1 * w1 * F, 2); need((size_t)B * hp * wp * F, 2);
    need((size_t)B * hp * wp * F2, 2); need(npix * F2, 2); need(npix * 32, 2); need(npix, 4);
    need((size_t)B * h4 * w4 * F, 2); need((size_t)B * h4 * w4 * 256, 2); need((size_t)B * h4 * w4 * 64, 2);
    for (int i = 0; i < 5; ++i) need((size_t)B * hp * wp * 128, 2);
    need((size_t)B * hp * wp * 16, 2);
    need((size_t)B * hp * wp * 64, 4); need((size_t)B * hp * wp * 64, 4);
    need(npix * 192, 2); need(npix * 96, 2);
    if (m->ensure_ws(bytes)) return 1;
    Arena a{m->ws, 0, m->ws_bytes};
    __half* Apatch = a.take<__half>((size_t)B * P * 768);
    __half* T = a.take<__half>((size_t)B * P * tcols);   // patch GEMM output, readout concat, reassemble-0 GEMM output
    float* X32 = a.take<float>((size_t)M * dim);
    __half* Hn = a.take<__half>((size_t)M * dim);
    __half* QKV = a.take<__half>((size_t)M * 3 * dim);
    __half* ATT = a.take<__half>((size_t)M * dim);
    __half* HID = a.take<__half>((size_t)M * 4 * dim);
    __half* D = a.take<__half>((size_t)M * dim);
    __half* FE[4];
    for (int i = 0; i < 4; ++i) FE[i] = a.take<__half>((size_t)M * dim);
    __half* Y = a.take<__half>((size_t)B * P * dim);
    __half* L1 = a.take<__half>((size_t)B * h1 * w1 * c0);
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
    __half* PATH[4];
    PATH[0] = a.take<__half>((size_t)B * h3 * w3 * F);
    PATH[1] = a.take<__half>((size_t)B * h2 * w2 * F);
    PATH[2] = a.take<__half>((size_t)B * h1 * w1 * F);
    PATH[3] = a.take<__half>((size_t)B * hp * wp * F);
    __half* O1 = a.take<__half>((size_t)B * hp * wp * F2);
    __half* O1u = a.take<__half>(npix * F2);
    __half* O2 = a.take<__half>(npix * 32);
    float* REL = a.take<float>(npix);
    __half* XB = a.take<__half>((size_t)B * h4 * w4 * F);
    __half* S1 = a.take<__half>((size_t)B * h4 * w4 * 256);
    __half* S2 = a.take<__half>((size_t)B * h4 * w4 * 64);
    __half* Ea = a.take<__half>((size_t)B * hp * wp * 128);
    __half* E[2] = {a.take<__half>((size_t)B * hp * wp * 128), a.take<__half>((size_t)B * hp * wp * 128)};
    __half* Yb = a.take<__half>((size_t)B * hp * wp * 128);
    __half* A1 = a.take<__half>((size_t)B * hp * wp * 128);
    __half* A2 = a.take<__half>((size_t)B * hp * wp * 16);
    float* BIN[2] = {a.take<float>((size_t)B * hp * wp * 64), a.take<float>((size_t)B * hp * wp * 64)};
    __half* CC = a.take<__half>(npix * 192);
    __half* G = a.take<__half>(npix * 96);

    // ---- encoder (timm beit.py Beit.forward_features with the MiDaS relative-position resample; hooks, no final norm)
    if (zoe_patch_im2col(st, x, B, H, W, Apatch)) return 1;
    if (da_linear(st, m, w.patch, Apatch, (long long)B * P, 768, T, ACT_NONE)) return 1;
    if (zoe_assemble_tokens(st, T, m->at<float>(w.cls), X32, B, P, dim)) return 1;
    const __half* pending = nullptr;
    int nf = 0;
    for (int i = 0; i < w.depth; ++i) {
        const ZoeBlockW& b = w.blocks[i];
        if (da_add_layernorm(st, X32, pending, m->at<float>(b.n1w), m->at<float>(b.n1b), Hn, M, dim)) return 1;
        if (da_linear(st, m, b.qkv, Hn, M, dim, QKV, ACT_NONE)) return 1;
        if (da_attention(st, QKV, ATT, B, N, w.heads, w.bias_dev + (size_t)i * bias_layer, w.ldb)) return 1;
        if (da_linear(st, m, b.proj, ATT, M, dim, D, ACT_NONE)) return 1;                 // gamma_1 folded
        if (da_add_layernorm(st, X32, D, m->at<float>(b.n2w), m->at<float>(b.n2b), Hn, M, dim)) return 1;
        if (da_linear(st, m, b.fc1, Hn, M, dim, HID, ACT_GELU)) return 1;
        if (da_linear(st, m, b.fc2, HID, M, 4 * dim, D, ACT_NONE)) return 1;              // gamma_2 folded
        pending = D;
        if (nf < 4 && i == w.hooks[nf]) {
            if (zoe_add_cast(st, X32, D, FE[nf], M * dim)) return 1;
            if (tap_copy(st, nf, FE[nf], (size_t)M * dim * 2)) return 1;                  // taps 0..3: hooked hidden states
            pending = nullptr;
            ++nf;
        }
    }
    // ---- DPT reassemble: ProjectReadout (cat(token, cls) -> Linear -> GELU), 1x1 conv, resize
    auto reassemble = [&](int i, __half* out, int out_mode, int cout) {
        if (zoe_readout_concat(st, FE[i], B, P, dim, T)) return 1;
        if (da_linear(st, m, w.readout[i], T, (long long)B * P, 2 * dim, Y, ACT_GELU)) return 1;
        ConvGemm g;
        g.A = Y; g.B = B; g.Hi = ph; g.Wi = pw; g.Ci = dim; g.Cin = dim; g.kind = CG_LINEAR_2D;
        g.Wt = m->at<__half>(w.reasm[i].w); g.N = w.reasm[i].N; g.bias = m->at<float>(w.reasm[i].b); g.act = ACT_NONE;
        g.out = out; g.ldo = out_mode == OUT_PIXSHUF2 ? cout : w.reasm[i].N; g.out_mode = out_mode; g.cout = cout;
        return conv_gemm(st, g);
    };
    if (reassemble(0, T, 0, 0)) return 1;                                   // [B*P][16*c0]   (T: the concat was consumed by the readout GEMM)
    if (da_depth_to_space4(st, T, B, ph, pw, c0, L1, c0)) return 1;         // [B][4ph][4pw][c0]
    if (reassemble(1, L2, OUT_PIXSHUF2, c1)) return 1;                      // [B][2ph][2pw][c1]
    if (reassemble(2, L3, 0, 0)) return 1;
    if (reassemble(3, L4lin, 0, 0)) return 1;
    if (da_im2col_s2(st, L4lin, B, h3, w3, c3, L4col)) return 1;
    if (da_linear(st, m, w.resize3, L4col, (long long)B * h4 * w4, 9 * c3, L4, ACT_NONE)) return 1;
    if (da_conv(st, m, w.rn[0], L1, B, h1, w1, c0, c0, R1, ACT_NONE, nullptr, true)) return 1;
    if (da_conv(st, m, w.rn[1], L2, B, h2, w2, c1, c1, R2, ACT_NONE, nullptr, true)) return 1;
    if (da_conv(st, m, w.rn[2], L3, B, h3, w3, c2, c2, R3, ACT_NONE, nullptr, true)) return 1;
    if (da_conv(st, m, w.rn[3], L4, B, h4, w4, c3, c3, R4, ACT_NONE, nullptr, true)) return 1;
    if (da_fusion(st, m, w.ref[3], F, R4, nullptr, B, h4, w4, h3, w3, t_relu, t_c1, t_sum, t_u, t_up, PATH[0])) return 1;
    if (da_fusion(st, m, w.ref[2], F, PATH[0], R3, B, h3, w3, h2, w2, t_relu, t_c1, t_sum, t_u, t_up, PATH[1])) return 1;
    if (da_fusion(st, m, w.ref[1], F, PATH[1], R2, B, h2, w2, h1, w1, t_relu, t_c1, t_sum, t_u, t_up, PATH[2])) return 1;
    if (da_fusion(st, m, w.ref[0], F, PATH[2], R1, B, h1, w1, hp, wp, t_relu, t_c1, t_sum, t_u, t_up, PATH[3])) return 1;
    for (int i = 0; i < 4; ++i)                                                             // taps 4..7: path_4 .. path_1 (NHWC)
        if (tap_copy(st, 4 + i, PATH[i], (size_t)B * (i == 0 ? h3 * w3 : i == 1 ? h2 * w2 : i == 2 ? h1 * w1 : hp * wp) * F * 2)) return 1;
    if (tap_copy(st, 8, R4, (size_t)B * h4 * w4 * F * 2)) return 1;                        // tap 8: bottleneck (layer4_rn)
    // ---- scratch.output_conv: relative depth + the 32-channel activation the bins head is conditioned on
    if (da_conv(st, m, w.oc1, PATH[3], B, hp, wp, F, F, O1, ACT_NONE, nullptr, true)) return 1;
    if (da_upsample_bilinear(st, O1, B, hp, wp, F2, O1u, H, W)) return 1;
    if (da_conv(st, m, w.oc2, O1u, B, H, W, F2, F2, O2, ACT_RELU, nullptr, true)) return 1;
    if (da_head_final(st, O2, (long long)npix, 32, m->at<float>(w.oc3w), w.oc3b, REL)) return 1;
    if (tap_copy(st, 9, O2, npix * 32 * 2)) return 1;                                       // tap 9: out_conv activation (NHWC, 32)
    if (tap_copy(st, 10, REL, npix * 4)) return 1;                                          // tap 10: relative depth fp32
    // ---- metric bins head (zoedepth_v1.py forward)
    auto c1x1 = [&](const Lin& l, const __half* A, int h, int ww, int cin, __half* out, int act) {
        return da_conv(st, m, l, A, B, h, ww, cin, cin, out, act, nullptr, false);
    };
    if (c1x1(w.conv2, R4, h4, w4, F, XB, ACT_NONE)) return 1;
    if (c1x1(w.seed1, XB, h4, w4, F, S1, ACT_RELU)) return 1;
    if (c1x1(w.seed2, S1, h4, w4, 256, S2, ACT_NONE)) return 1;
    if (zoe_softplus(st, S2, BIN[0], (long long)B * h4 * w4 * 64)) return 1;
    if (c1x1(w.sproj1, XB, h4, w4, F, Ea, ACT_RELU)) return 1;
    if (c1x1(w.sproj2, Ea, h4, w4, 128, E[0], ACT_NONE)) return 1;
    const int lh[4] = {h3, h2, h1, hp}, lw[4] = {w3, w2, w1, wp};
    int prev_h = h4, prev_w = w4, cur = 0;
    for (int i = 0; i < 4; ++i) {
        const int nh = lh[i], nw = lw[i], nxt = cur ^ 1;
        if (c1x1(w.proj1[i], PATH[i], nh, nw, F, Ea, ACT_RELU)) return 1;
        if (c1x1(w.proj2[i], Ea, nh, nw, 128, E[nxt], ACT_NONE)) return 1;
        if (zoe_add_upsampled(st, E[nxt], E[cur], B, prev_h, prev_w, 128, nh, nw, Yb)) return 1;
        if (c1x1(w.att1[i], Yb, nh, nw, 128, A1, ACT_RELU)) return 1;
        if (c1x1(w.att2[i], A1, nh, nw, 128, A2, ACT_NONE)) return 1;
        if (zoe_attractor(st, A2, 16, w.n_att[i], BIN[cur], B, prev_h, prev_w, nh, nw, BIN[nxt])) return 1;
        prev_h = nh; prev_w = nw; cur = nxt;
        if (tap_copy(st, 11 + i, BIN[cur], (size_t)B * nh * nw * 64 * 4)) return 1;         // taps 11..14: bin centres fp32 [pix][64]
    }
    if (zoe_clb_concat(st, O2, REL, E[cur], B, hp, wp, H, W, CC)) return 1;
    if (da_linear(st, m, w.clb1, CC, (long long)npix, 192, G, ACT_GELU)) return 1;
    return zoe_clb_final(st, G, 96, m->at<float>(w.clb2w), m->at<float>(w.clb2b), BIN[cur], B, hp, wp, H, W, depth);
}

}  // namespace nb200
