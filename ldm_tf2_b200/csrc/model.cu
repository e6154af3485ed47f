// Model graphs of the sampling path on the B200 engine.  See model.h.
#include "model.h"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <map>
#include <algorithm>

namespace ldm {

static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <typename T>
T* Model::dev_alloc(size_t n, bool zero) {
  if (eng.device < 0) return nullptr;   // describe-only model (ldm_create with device -1)
  void* p = nullptr;
  CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  // zero on the engine's own (non-blocking) stream: a legacy-stream memset is unordered against it
  if (zero) CUDA_CHECK(cudaMemsetAsync(p, 0, std::max<size_t>(n, 1) * sizeof(T), eng.stream));
  owned_.push_back(p);
  return reinterpret_cast<T*>(p);
}

// Frees a dev_alloc'ed buffer that is being replaced (after the stream has drained: kernels in flight
// may still read it).
void Model::dev_free(void* p) {
  if (!p) return;
  auto it = std::find(owned_.begin(), owned_.end(), p);
  if (it == owned_.end()) return;
  eng.sync();
  owned_.erase(it);
  cudaFree(p);
}


// =====================================================================================
// Construction: slots in flat Keras order + packed destinations
// =====================================================================================
struct Builder {
  Model& m;
  int model;
  std::vector<Slot>& s;
  explicit Builder(Model& mm, int which) : m(mm), model(which), s(mm.slots[which]) {}

  Slot* f32(const std::string& name, std::vector<int> shape) {
    Slot sl;
    sl.name = name;
    sl.shape = std::move(shape);
    sl.kind = Slot::F32;
    s.push_back(sl);
    return &s.back();
  }
  // kernel viewed as [k, n]; packed to dst[(row0+perm(n))*ld + col0 + k]
  Slot* pack(const std::string& name, std::vector<int> shape, int k, int n, bf16* dst, long long ld, int row0,
             int col0, int geglu_half = 0) {
    Slot sl;
    sl.name = name;
    sl.shape = std::move(shape);
    sl.kind = Slot::PACK;
    sl.dst = dst; sl.ld = ld; sl.row0 = row0; sl.col0 = col0; sl.k = k; sl.n = n; sl.geglu_half = geglu_half;
    s.push_back(sl);
    return &s.back();
  }
  Slot* f32mat(const std::string& name, std::vector<int> shape, int k, int n, float* dst, long long ld, int col0) {
    Slot sl;
    sl.name = name;
    sl.shape = std::move(shape);
    sl.kind = Slot::F32MAT;
    sl.k = k; sl.n = n; sl.f32_dst = dst; sl.f32_ld = ld; sl.f32_col0 = col0;
    s.push_back(sl);
    return &s.back();
  }
  LinW lin(int n, int k) {
    LinW w;
    w.n = n; w.k = k; w.ld = k;
    LDM_CHECK(k % 8 == 0, "linear K=%d must be a multiple of 8 (16-byte TMA rows)", k);
    w.wt = m.dev_alloc<bf16>((size_t)n * k, true);
    return w;
  }
  // up-conv kernel slot `k` [3,3,c,c]: second packing as four phase-collapsed 2x2 kernels
  LinW up_phase(Slot* k, int c) {
    LinW w;
    w.n = c; w.k = 4 * c; w.ld = 4 * c;
    w.wt = m.dev_alloc<bf16>((size_t)4 * c * 4 * c, true);
    k->up_dst = w.wt; k->up_cin = c; k->up_cout = c;
    return w;
  }
  GNW gnw(const std::string& p, int c, float eps) {
    GNW g; g.c = c; g.eps = eps;
    g.gamma = f32(p + "/gamma", {c});
    g.beta = f32(p + "/beta", {c});
    return g;
  }
  LNW lnw(const std::string& p, int c) {
    LNW g; g.c = c;
    g.gamma = f32(p + "/gamma", {c});
    g.beta = f32(p + "/beta", {c});
    return g;
  }
  // Dense [k,n] + bias
  LinW dense(const std::string& p, int k, int n, bool bias = true) {
    LinW w = lin(n, k);
    pack(p + "/kernel", {k, n}, k, n, w.wt, w.ld, 0, 0);
    if (bias) w.bias = f32(p + "/bias", {n});
    return w;
  }
  // ResidualBlock (unet.py:368-380 / autoencoder.py:13-41) in Keras weight order
  ResW res(const std::string& p, int cin, int cout, int temb_dim, bool shortcut, bool ae, int* temb_cols,
           float* tproj, int tproj_ld) {
    ResW r;
    r.cin = cin; r.cout = cout; r.shortcut = shortcut;
    const char* n_gn1 = ae ? "/_group_norm1" : "/_group_norm_1";
    const char* n_c1 = ae ? "/_conv1" : "/_conv2d_1";
    const char* n_gn2 = ae ? "/_group_norm2" : "/_group_norm_2";
    const char* n_c2 = ae ? "/_conv2" : "/_conv2d_2";
    const float eps = ae ? 1e-6f : 1e-5f;
    r.gn1 = gnw(p + n_gn1, cin, eps);
    r.conv1 = lin(cout, 9 * cin);
    pack(p + n_c1 + "/kernel", {3, 3, cin, cout}, 9 * cin, cout, r.conv1.wt, r.conv1.ld, 0, 0);
    r.conv1.bias = f32(p + n_c1 + "/bias", {cout});
    if (temb_dim) {
      r.temb_off = *temb_cols;
      f32mat(p + "/_dense/kernel", {temb_dim, cout}, temb_dim, cout, tproj, tproj_ld, r.temb_off);
      Slot* b = f32(p + "/_dense/bias", {cout});
      tproj_bias.push_back({b, r.temb_off});
      *temb_cols += cout;
    }
    r.gn2 = gnw(p + n_gn2, cout, eps);
    r.conv2 = lin(cout, 9 * cout + (shortcut ? cin : 0));
    pack(p + n_c2 + "/kernel", {3, 3, cout, cout}, 9 * cout, cout, r.conv2.wt, r.conv2.ld, 0, 0);
    r.conv2.bias = f32(p + n_c2 + "/bias", {cout});
    if (shortcut) {
      pack(p + "/_shortcut/kernel", {cin, cout}, cin, cout, r.conv2.wt, r.conv2.ld, 0, 9 * cout);
      r.sc_bias = f32(p + "/_shortcut/bias", {cout});
    }
    return r;
  }
  std::vector<std::pair<Slot*, int>> tproj_bias;

  // CrossAttention (unet.py:248-267): q,k,v split Projections (no bias), out merge (+bias)
  AttnW attn(const std::string& p, int cq, int ckv, int heads, int d, int cout, bool self) {
    AttnW a;
    a.heads = heads; a.d = d;
    const int inner = heads * d;
    if (self) {
      a.qkv = lin(3 * inner, cq);
      a.sq = pack(p + "/_dense_layer_query/kernel", {cq, heads, d}, cq, inner, a.qkv.wt, a.qkv.ld, 0, 0);
      a.sk = pack(p + "/_dense_layer_key/kernel", {cq, heads, d}, cq, inner, a.qkv.wt, a.qkv.ld, inner, 0);
      a.sv = pack(p + "/_dense_layer_value/kernel", {cq, heads, d}, cq, inner, a.qkv.wt, a.qkv.ld, 2 * inner, 0);
    } else {
      a.qkv = lin(inner, cq);
      a.kv = lin(2 * inner, ckv);
      a.sq = pack(p + "/_dense_layer_query/kernel", {cq, heads, d}, cq, inner, a.qkv.wt, a.qkv.ld, 0, 0);
      pack(p + "/_dense_layer_key/kernel", {ckv, heads, d}, ckv, inner, a.kv.wt, a.kv.ld, 0, 0);
      pack(p + "/_dense_layer_value/kernel", {ckv, heads, d}, ckv, inner, a.kv.wt, a.kv.ld, inner, 0);
    }
    a.out = lin(cout, inner);
    pack(p + "/_dense_layer_output/kernel", {heads, d, cout}, inner, cout, a.out.wt, a.out.ld, 0, 0);
    a.out.bias = f32(p + "/_dense_layer_output/bias", {cout});
    return a;
  }
  // SpatialTransformer (unet.py:341-354), 26 tensors, GroupNorm last
  void st(STW& t, const std::string& p, int c, int heads, int d, int ctx) {
    t.c = c; t.d = d;
    t.d1 = dense(p + "/_dense1", c, c);
    t.a1 = attn(p + "/_block/_att_layer1", c, c, heads, d, c, true);
    t.a2 = attn(p + "/_block/_att_layer2", c, ctx, heads, d, c, false);
    // GEGLU Dense c -> 8c, rows permuted per tile so value/gate columns share a tile
    t.geglu = lin(8 * c, c);
    int bn = 256;
    while ((8 * c) % bn) bn -= 64;   // value and gate halves are processed in 32-column chunks
    LDM_CHECK(bn >= 64, "GEGLU width %d has no tile that is a multiple of 64", 8 * c);
    t.geglu_bn = bn;
    t.geglu_k = pack(p + "/_block/_ffn_layer/_geglu_layer/_dense_layer/kernel", {c, 8 * c}, c, 8 * c, t.geglu.wt,
                     t.geglu.ld, 0, 0, bn / 2);
    t.geglu.bias = f32(p + "/_block/_ffn_layer/_geglu_layer/_dense_layer/bias", {8 * c});
    t.ff = dense(p + "/_block/_ffn_layer/_dense_layer", 4 * c, c);
    t.ln1 = lnw(p + "/_block/_layernorm1", c);
    t.ln2 = lnw(p + "/_block/_layernorm2", c);
    t.ln3 = lnw(p + "/_block/_layernorm3", c);
    t.d2 = dense(p + "/_dense2", c, c);
    t.gn = gnw(p + "/_groupnorm", c, 1e-6f);
    // LayerNorm1 -> q|k|v, LayerNorm2 -> q of the cross attention, LayerNorm3 -> GEGLU (unet.py:309-313) are
    // folded into those linears: the kernels keep their fp32 copies until finalize_weights has gamma / beta
    for (Slot* k : {t.a1.sq, t.a1.sk, t.a1.sv}) { k->keep = true; m.folds_.push_back({k, &t.ln1, &t.a1.qkv, nullptr, nullptr}); }
    t.a2.sq->keep = true;
    m.folds_.push_back({t.a2.sq, &t.ln2, &t.a2.qkv, nullptr, nullptr});
    t.geglu_k->keep = true;
    m.folds_.push_back({t.geglu_k, &t.ln3, &t.geglu, t.geglu.bias, &t});
    for (LinW* w : {&t.a1.qkv, &t.a2.qkv, &t.geglu}) {
      w->ln_cs = m.dev_alloc<float>(w->n, true);
      w->ln_bias = m.dev_alloc<float>(w->n, true);
    }
  }
  // AE AttentionBlock (autoencoder.py:61-72): GN, then q,k,v,out Dense with bias
  void ae_attn(AEAttnW& a, const std::string& p, int c) {
    a.c = c;
    a.gn = gnw(p + "/_group_norm", c, 1e-6f);
    a.qkv = lin(3 * c, c);
    a.qkv_bias = m.dev_alloc<float>(3 * c, true);
    const char* nm[3] = {"/_dense_query", "/_dense_key", "/_dense_value"};
    for (int i = 0; i < 3; ++i) {
      pack(p + nm[i] + "/kernel", {c, c}, c, c, a.qkv.wt, a.qkv.ld, i * c, 0);
      Slot* b = f32(p + nm[i] + "/bias", {c});
      concat_bias.push_back({b, a.qkv_bias + i * c});
    }
    a.out = dense(p + "/_dense_output", c, c);
  }
  std::vector<std::pair<Slot*, float*>> concat_bias;
};

Model::Model(const ModelConfig& c, int device) : cfg(c), eng(device) {
  eng.fp16 = c.precision >= 1 ? 1 : 0;   // validation mode (2): the parts that stay 16-bit use fp16
  // Residual stream format at block boundaries.  Default: 16 bit (fp32 only inside accumulators and statistics),
  // i.e. every GEMM epilogue writes one 16-bit tensor and GroupNorm reads 2 bytes per element; LDM_B200_STREAM=fp32
  // keeps an fp32 stream with a 16-bit shadow (round 1's layout) for accuracy comparisons.
  {
    const char* e = getenv("LDM_B200_STREAM");
    stream16_ = !(e && !strcmp(e, "fp32"));
  }
  build();
}

Model::~Model() {
  if (comm_) comm_destroy();
  if (step_graph_) cudaGraphExecDestroy(step_graph_);
  for (auto& v : slots)
    for (auto& s : v)
      if (s.f32) cudaFree(s.f32);
  for (void* p : owned_) cudaFree(p);
  for (void* p : stage_ptr_) if (p) cudaFree(p);
  if (ev0_) { cudaEventDestroy(ev0_); cudaEventDestroy(ev1_); }
}

void Model::build() {
  // ---------------- text transformer (transformer.py:218-252; flat order SURVEY A.3)
  {
    Builder b(*this, 0);
    const int D = cfg.text_hidden, H = cfg.text_heads, S = cfg.text_head_dim, F = cfg.text_filter;
    slots[0].reserve((size_t)cfg.text_layers * 13 + 4);
    text_layers_.resize(cfg.text_layers);
    for (int i = 0; i < cfg.text_layers; ++i) {
      const std::string p = "transformer/_encoder/_stack/" + std::to_string(i);
      TextLayer& L = text_layers_[i];
      L.attn = b.attn(p + "/_mha", D, D, H, S, D, true);
      L.ln_mha = b.lnw(p + "/_layernorm_mha", D);
      L.f1 = b.dense(p + "/_ffn/_dense_layer_filter", D, F);
      L.f2 = b.dense(p + "/_ffn/_dense_layer_output", F, D);
      L.ln_ffn = b.lnw(p + "/_layernorm_ffn", D);
    }
    text_ln_ = b.lnw("transformer/_encoder/_layernorm", D);
    tok_emb_ = b.f32("transformer/_embedding_layer/embeddings", {cfg.vocab_size, D});
    pos_emb_ = b.f32("transformer/_positional_embedding_layer/embeddings", {cfg.max_seq_len, D});
  }
  // ---------------- unet (unet.py:51-116)
  {
    Builder b(*this, 1);
    slots[1].reserve(4096);
    const int mc = cfg.model_channels, td = 4 * mc, heads = cfg.num_heads, ctx = cfg.context_dim;
    const int L = cfg.num_mult, nb = cfg.num_blocks;
    // total columns of the stacked time projections
    int sumc = 0;
    {
      for (int i = 0; i < L; ++i) sumc += nb * mc * cfg.channel_mult[i];
      sumc += 2 * mc * cfg.channel_mult[L - 1];
      for (int i = L - 1; i >= 0; --i) sumc += (nb + 1) * mc * cfg.channel_mult[i];
    }
    tproj_w_ = dev_alloc<float>((size_t)td * sumc, true);
    tproj_bias_ = dev_alloc<float>(sumc, true);
    time1_w_ = dev_alloc<float>((size_t)mc * td, true);
    time2_w_ = dev_alloc<float>((size_t)td * td, true);
    tproj_cols_ = sumc;
    int tcols = 0;
    conv_in_k_ = b.f32("unet/_conv_in/kernel", {3, 3, 4, mc});
    conv_in_b_ = b.f32("unet/_conv_in/bias", {mc});
    b.f32mat("unet/_time_dense1/kernel", {mc, td}, mc, td, time1_w_, td, 0);
    time1_b_ = b.f32("unet/_time_dense1/bias", {td});
    b.f32mat("unet/_time_dense2/kernel", {td, td}, td, td, time2_w_, td, 0);
    time2_b_ = b.f32("unet/_time_dense2/bias", {td});
    std::vector<int> chans{mc};
    int ch = mc;
    in_blocks_.reserve(64);
    out_blocks_.reserve(64);
    for (int i = 0; i < L; ++i) {
      const int m = cfg.channel_mult[i];
      for (int j = 0; j < nb; ++j) {
        const std::string p = "unet/_input_blocks/" + std::to_string(in_blocks_.size());
        in_blocks_.emplace_back();
        UNetBlock& blk = in_blocks_.back();
        blk.kind = 0; blk.cin = ch; blk.cout = mc * m;
        blk.res = b.res(p + "/_residual", ch, mc * m, td, ch != mc * m, false, &tcols, tproj_w_, sumc);
        blk.has_st = i < L - 1;
        if (blk.has_st) b.st(blk.st, p + "/_spatial_transformer", mc * m, heads, cfg.head_base * m, ctx);
        ch = mc * m;
        chans.push_back(ch);
      }
      if (i < L - 1) {
        const std::string p = "unet/_input_blocks/" + std::to_string(in_blocks_.size());
        in_blocks_.emplace_back();
        UNetBlock& blk = in_blocks_.back();
        blk.kind = 1; blk.cin = blk.cout = ch;
        blk.resample = b.lin(ch, 9 * ch);
        b.pack(p + "/_downsample/_conv/kernel", {3, 3, ch, ch}, 9 * ch, ch, blk.resample.wt, blk.resample.ld, 0, 0);
        blk.resample.bias = b.f32(p + "/_downsample/_conv/bias", {ch});
        chans.push_back(ch);
      }
    }
    mid_res1_ = b.res("unet/_middle_block/_residual1", ch, ch, td, false, false, &tcols, tproj_w_, sumc);
    b.st(mid_st_, "unet/_middle_block/_spatial_transformer", ch, heads, cfg.head_base * cfg.channel_mult[L - 1], ctx);
    mid_res2_ = b.res("unet/_middle_block/_residual2", ch, ch, td, false, false, &tcols, tproj_w_, sumc);
    for (int i = L - 1; i >= 0; --i) {
      const int m = cfg.channel_mult[i];
      for (int j = 0; j <= nb; ++j) {
        const std::string p = "unet/_output_blocks/" + std::to_string(out_blocks_.size());
        const int skip = chans.back();
        chans.pop_back();
        out_blocks_.emplace_back();
        UNetBlock& blk = out_blocks_.back();
        blk.kind = 2; blk.cin = ch + skip; blk.cout = mc * m;
        blk.res = b.res(p + "/_residual", ch + skip, mc * m, td, true, false, &tcols, tproj_w_, sumc);
        blk.has_st = i < L - 1;
        if (blk.has_st) b.st(blk.st, p + "/_spatial_transformer", mc * m, heads, cfg.head_base * m, ctx);
        blk.has_up = (i > 0 && j == nb);
        ch = mc * m;
        if (blk.has_up) {
          blk.resample = b.lin(ch, 9 * ch);
          blk.up_phase = b.up_phase(b.pack(p + "/_upsample/_conv/kernel", {3, 3, ch, ch}, 9 * ch, ch, blk.resample.wt,
                                           blk.resample.ld, 0, 0), ch);
          blk.resample.bias = b.f32(p + "/_upsample/_conv/bias", {ch});
        }
      }
    }
    LDM_CHECK(tcols == sumc, "time projection column count mismatch %d vs %d", tcols, sumc);
    out_gn_ = b.gnw("unet/_groupnorm", mc, 1e-5f);
    conv_out_ = b.lin(cfg.out_channels, 9 * mc);
    b.pack("unet/_conv_out/kernel", {3, 3, mc, cfg.out_channels}, 9 * mc, cfg.out_channels, conv_out_.wt,
           conv_out_.ld, 0, 0);
    conv_out_.bias = b.f32("unet/_conv_out/bias", {cfg.out_channels});
    // remember where each time-Dense bias goes
    tproj_bias_slots_.assign(b.tproj_bias.begin(), b.tproj_bias.end());
    for (auto& blk : in_blocks_) if (blk.has_st) all_st_.push_back(&blk.st);
    all_st_.push_back(&mid_st_);
    for (auto& blk : out_blocks_) if (blk.has_st) all_st_.push_back(&blk.st);
  }
  // ---------------- autoencoder decode side (autoencoder.py:252-290,331-347,408-421)
  {
    Builder b(*this, 2);
    slots[2].reserve(1024);
    const int z = cfg.latent_channels, chn = cfg.ae_channels, L = cfg.ae_num_mult, nb = cfg.ae_num_blocks;
    std::vector<int> chans(L);
    for (int i = 0; i < L; ++i) chans[i] = chn * cfg.ae_mult[i];
    const int top = chans[L - 1];
    if (cfg.ae_kind == 1) codebook_ = b.f32("autoencoder/_quantize/kernel", {cfg.vq_vocab, z});
    pq_k_ = b.f32("autoencoder/_post_quant_conv/kernel", {z, z});
    pq_b_ = b.f32("autoencoder/_post_quant_conv/bias", {z});
    const std::string d = "autoencoder/_decoder";
    ae_conv_in_k_ = b.f32(d + "/_conv_in/kernel", {3, 3, z, top});
    ae_conv_in_b_ = b.f32(d + "/_conv_in/bias", {top});
    ae_mid1_ = b.res(d + "/_middle/_residual1", top, top, 0, false, true, nullptr, nullptr, 0);
    b.ae_attn(ae_mid_attn_, d + "/_middle/_attention", top);
    ae_mid2_ = b.res(d + "/_middle/_residual2", top, top, 0, false, true, nullptr, nullptr, 0);
    ae_up_.reserve(64);
    int hw = cfg.ae_build_hw, cur = top, idx = 0;
    for (int i = L - 1; i >= 0; --i) {
      for (int j = 0; j <= nb; ++j) {
        const std::string p = d + "/_up/" + std::to_string(idx++);
        ae_up_.emplace_back();
        AEStage& s = ae_up_.back();
        s.kind = 0; s.hw = hw;
        s.res = b.res(p + "/_residual", cur, chans[i], 0, cur != chans[i], true, nullptr, nullptr, 0);
        s.attn = false;
        if (cfg.ae_kind == 1)
          for (int k = 0; k < cfg.ae_num_attn_res; ++k) s.attn |= (cfg.ae_attn_res[k] == hw);
        if (s.attn) b.ae_attn(s.at, p + "/_attention", chans[i]);
        cur = chans[i];
      }
      if (i > 0) {
        const std::string p = d + "/_up/" + std::to_string(idx++);
        ae_up_.emplace_back();
        AEStage& s = ae_up_.back();
        s.kind = 1; s.c = cur; s.hw = hw;
        s.up = b.lin(cur, 9 * cur);
        s.up_phase = b.up_phase(b.pack(p + "/_conv/kernel", {3, 3, cur, cur}, 9 * cur, cur, s.up.wt, s.up.ld, 0, 0), cur);
        s.up.bias = b.f32(p + "/_conv/bias", {cur});
        hw *= 2;
      }
    }
    ae_out_gn_ = b.gnw(d + "/_group_norm", chans[0], 1e-6f);
    ae_conv_out_ = b.lin(3, 9 * chans[0]);
    b.pack(d + "/_conv_out/kernel", {3, 3, chans[0], 3}, 9 * chans[0], 3, ae_conv_out_.wt, ae_conv_out_.ld, 0, 0);
    ae_conv_out_.bias = b.f32(d + "/_conv_out/bias", {3});
    ae_concat_bias_.assign(b.concat_bias.begin(), b.concat_bias.end());
  }
  // ---------------- autoencoder encode side (autoencoder.py:198-249,322-331,395-405): flat Keras order of a
  // model on which only encode() was called (oracle.ae_encoder_spec)
  {
    Builder b(*this, 3);
    slots[3].reserve(1024);
    const int z = cfg.latent_channels * (cfg.ae_kind == 0 ? 2 : 1), chn = cfg.ae_channels, L = cfg.ae_num_mult,
              nb = cfg.ae_num_blocks;
    enc_z_ = z;
    const std::string e = "autoencoder/_encoder";
    enc_conv_in_k_ = b.f32(e + "/_conv_in/kernel", {3, 3, 3, chn});
    enc_conv_in_b_ = b.f32(e + "/_conv_in/bias", {chn});
    enc_down_.reserve(64);
    int hw = cfg.ae_build_hw << (L - 1), cur = chn, idx = 0;   // image side the checkpoint's Encoder was built at
    for (int i = 0; i < L; ++i) {
      const int co = chn * cfg.ae_mult[i];
      for (int j = 0; j < nb; ++j) {
        const std::string p = e + "/_down/" + std::to_string(idx++);
        enc_down_.emplace_back();
        EncStage& s = enc_down_.back();
        s.kind = 0; s.hw = hw;
        s.res = b.res(p + "/_residual", cur, co, 0, cur != co, true, nullptr, nullptr, 0);
        s.attn = false;
        if (cfg.ae_kind == 1)   // AutoencoderKL builds its Encoder with attention_resolutions=() (autoencoder.py:326)
          for (int k = 0; k < cfg.ae_num_attn_res; ++k) s.attn |= (cfg.ae_attn_res[k] == hw);
        if (s.attn) b.ae_attn(s.at, p + "/_attention", co);
        cur = co;
      }
      if (i < L - 1) {
        const std::string p = e + "/_down/" + std::to_string(idx++);
        enc_down_.emplace_back();
        EncStage& s = enc_down_.back();
        s.kind = 1; s.c = cur; s.hw = hw;
        s.down = b.lin(cur, 9 * cur);
        b.pack(p + "/_conv/kernel", {3, 3, cur, cur}, 9 * cur, cur, s.down.wt, s.down.ld, 0, 0);
        s.down.bias = b.f32(p + "/_conv/bias", {cur});
        hw /= 2;
      }
    }
    enc_mid1_ = b.res(e + "/_middle/_residual1", cur, cur, 0, false, true, nullptr, nullptr, 0);
    b.ae_attn(enc_mid_attn_, e + "/_middle/_attention", cur);
    enc_mid2_ = b.res(e + "/_middle/_residual2", cur, cur, 0, false, true, nullptr, nullptr, 0);
    enc_out_gn_ = b.gnw(e + "/_group_norm", cur, 1e-6f);
    enc_conv_out_ = b.lin(z, 9 * cur);
    b.pack(e + "/_conv_out/kernel", {3, 3, cur, z}, 9 * cur, z, enc_conv_out_.wt, enc_conv_out_.ld, 0, 0);
    enc_conv_out_.bias = b.f32(e + "/_conv_out/bias", {z});
    quant_k_ = b.f32("autoencoder/_quant_conv/kernel", {z, z});
    quant_b_ = b.f32("autoencoder/_quant_conv/bias", {z});
    for (auto& pr : b.concat_bias) ae_concat_bias_.push_back(pr);   // q|k|v biases of the encoder's attention blocks
    enc_concat_from_ = ae_concat_bias_.size() - b.concat_bias.size();
  }
  // validation mode walks the raw fp32 kernels of the unet (validate.cu)
  if (cfg.precision == 2)
    for (int mdl = 0; mdl < NUM_MODELS; ++mdl)
      for (auto& s : slots[mdl])
        if (s.kind == Slot::PACK) s.keep = true;
  step_dev_ = dev_alloc<int>(1, true);
  sat_dev_ = dev_alloc<unsigned long long>(1, true);
}

// =====================================================================================
// Weights
// =====================================================================================
void Model::set_weight(int model, int index, const float* src, const int* shape, int ndim) {
  LDM_CHECK(model >= 0 && model < NUM_MODELS, "set_weight: model %d", model);
  LDM_CHECK(index >= 0 && index < (int)slots[model].size(), "set_weight: index %d out of range (model %d has %d)",
            index, model, (int)slots[model].size());
  Slot& s = slots[model][index];
  bool ok = (int)s.shape.size() == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = s.shape[i] == shape[i];
  if (!ok) {
    std::string want, got;
    for (int d : s.shape) want += std::to_string(d) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    throw Error(fmt("set_weight: %s (model %d index %d) expects shape [%s] got [%s]", s.name.c_str(), model, index,
                    want.c_str(), got.c_str()));
  }
  CUDA_CHECK(cudaSetDevice(eng.device));
  const size_t n = s.numel();
  float* dev = nullptr;
  CUDA_CHECK(cudaMalloc(&dev, n * sizeof(float)));
  CUDA_CHECK(cudaMemcpyAsync(dev, src, n * sizeof(float), cudaMemcpyDefault, eng.stream));
  if (s.kind == Slot::F32) {
    // keep the address of an already installed tensor (captured graphs and hoisted tables hold it)
    if (s.f32) {
      CUDA_CHECK(cudaMemcpyAsync(s.f32, dev, n * sizeof(float), cudaMemcpyDeviceToDevice, eng.stream));
      eng.sync();
      cudaFree(dev);
    } else {
      eng.sync();
      s.f32 = dev;
    }
  } else if (s.kind == Slot::F32MAT) {
    CUDA_CHECK(cudaMemcpy2DAsync(s.f32_dst + s.f32_col0, s.f32_ld * sizeof(float), dev, s.n * sizeof(float),
                                 s.n * sizeof(float), s.k, cudaMemcpyDeviceToDevice, eng.stream));
    eng.sync();
    cudaFree(dev);
  } else {
    launch_pack_weight(dev, s.k, s.n, s.dst + s.col0, s.ld, s.row0, s.geglu_half, eng.fp16, eng.stream);
    if (s.up_dst) launch_pack_upconv_phase(dev, s.up_cin, s.up_cout, s.up_dst, eng.fp16, eng.stream);
    eng.sync();
    if (s.keep) {   // re-packed with the LayerNorm gamma by finalize_weights (apply_folds)
      if (s.f32) cudaFree(s.f32);
      s.f32 = dev;
    } else {
      cudaFree(dev);
    }
  }
  s.set = true;
  finalized = false;
  // anything derived from the old weights is stale: the captured step, the hoisted time-embedding table
  // (recomputed by finalize_weights) and the hoisted context K / V^T (set_context must run again)
  invalidate_graph();
  if (model == 1) { ctx_rows_ = 0; sampler_stale_ = true; }
}

void Model::finalize_weights() {
  CUDA_CHECK(cudaSetDevice(eng.device));
  // Only models whose weights were all provided are usable; partially set models are an error.
  for (int mdl = 0; mdl < NUM_MODELS; ++mdl) {
    int nset = 0;
    for (auto& s : slots[mdl]) nset += s.set ? 1 : 0;
    model_ready_[mdl] = nset == (int)slots[mdl].size();
    if (nset != 0 && !model_ready_[mdl]) {
      for (auto& s : slots[mdl])
        if (!s.set) throw Error(fmt("finalize_weights: model %d is missing %s (%d of %d set)", mdl, s.name.c_str(),
                                    nset, (int)slots[mdl].size()));
    }
  }
  if (model_ready_[1]) {
    for (auto& pr : tproj_bias_slots_)
      CUDA_CHECK(cudaMemcpyAsync(tproj_bias_ + pr.second, pr.first->f32, pr.first->numel() * sizeof(float),
                                 cudaMemcpyDeviceToDevice, eng.stream));
    apply_folds();
  }
  for (size_t i = 0; i < ae_concat_bias_.size(); ++i) {
    if (!model_ready_[i < enc_concat_from_ ? 2 : 3]) continue;
    auto& pr = ae_concat_bias_[i];
    CUDA_CHECK(cudaMemcpyAsync(pr.second, pr.first->f32, pr.first->numel() * sizeof(float),
                               cudaMemcpyDeviceToDevice, eng.stream));
  }
  eng.sync();
  finalized = true;
  if (model_ready_[1] && S_ > 0 && sampler_stale_) {
    compute_temb_table(ddim_t_.data(), S_, sampler_temb_);   // per-step time projections of the NEW weights
    sampler_stale_ = false;
  }
}

// LayerNorm folded into the linear that consumes it (unet.py:309-313):
//   LN(y) W + b = rstd_r (y (gamma o W) - mean_r 1^T (gamma o W)) + (beta W + b)
// so the GEMM runs on the RAW 16-bit rows y with weights gamma o W; its epilogue applies the per-row
// (mean, rstd) -- taken from the producer's epilogue (GemmOp::rs_out) -- with ln_cs = column sums of the
// packed weights and ln_bias = beta W + b.  Saves the LayerNorm pass and its normalised copy of y.
void Model::apply_folds() {
  std::vector<float> tmp_h;
  for (const Fold& f : folds_) {
    Slot& k = *f.kernel;
    LDM_CHECK(k.f32 && f.ln->gamma->f32 && f.ln->beta->f32, "apply_folds: %s not resident", k.name.c_str());
    launch_pack_weight(k.f32, k.k, k.n, k.dst + k.col0, k.ld, k.row0, k.geglu_half, eng.fp16, eng.stream, f.ln->gamma->f32);
    // beta W (+ b) in natural column order
    float* tmp = static_cast<float*>(stage(ST_A, (size_t)k.n * sizeof(float)));
    launch_small_dense_f32(f.ln->beta->f32, k.f32, f.bias_src ? f.bias_src->f32 : nullptr, 1, k.k, k.n, 0, 0, tmp, eng.stream);
    if (f.geglu_of) {
      // permute like the packed weight rows: value / gate columns of a tile side by side
      STW* st = f.geglu_of;
      const int n = k.n, nh = n / 2, half = st->geglu_bn / 2;
      tmp_h.resize(n);
      std::vector<float> pb(n);
      CUDA_CHECK(cudaMemcpyAsync(tmp_h.data(), tmp, n * sizeof(float), cudaMemcpyDeviceToHost, eng.stream));
      eng.sync();
      for (int c = 0; c < n; ++c) {
        const int j2 = c < nh ? c : c - nh;
        pb[(j2 / half) * (2 * half) + (c < nh ? 0 : half) + j2 % half] = tmp_h[c];
      }
      CUDA_CHECK(cudaMemcpyAsync(f.lin->ln_bias, pb.data(), n * sizeof(float), cudaMemcpyHostToDevice, eng.stream));
      eng.sync();
      st->geglu_bias_perm = f.lin->ln_bias;
    } else {
      CUDA_CHECK(cudaMemcpyAsync(f.lin->ln_bias + k.row0, tmp, (size_t)k.n * sizeof(float), cudaMemcpyDeviceToDevice, eng.stream));
    }
    eng.sync();
  }
  // column sums of the packed matrices, as the tensor cores see them (16-bit values, packed row order)
  for (const Fold& f : folds_)
    launch_rowsum16(f.lin->wt, f.lin->ld, f.kernel->row0, f.kernel->n, f.lin->k, f.lin->ln_cs + f.kernel->row0, eng.fp16, eng.stream);
  eng.sync();
}

void Model::ensure_arena(size_t bytes) {
  // called right after a dry pass: pool_off_ holds the statistic bytes it used
  if (pool_off_ > pool_need_) pool_need_ = pool_off_;
  if (pool_need_ > pool_cap_) {
    dev_free(pool_);
    pool_ = dev_alloc<uint8_t>(pool_need_, true);
    pool_cap_ = pool_need_;
    invalidate_graph();   // a captured step holds the old pool address
  }
  bytes += (64u << 20);
  if (eng.arena.capacity() >= bytes) return;
  eng.sync();
  CUDA_CHECK(cudaDeviceSynchronize());
  if (step_graph_) { cudaGraphExecDestroy(step_graph_); step_graph_ = nullptr; }
  eng.arena.init(bytes);
}

// =====================================================================================
// Ops
// =====================================================================================
void Model::tap(const std::string& name, const Act& a) {
  if (eng.dry || taps.empty()) return;
  auto it = taps.find(name);
  if (it == taps.end()) return;
  LDM_CHECK(it->second.second == (size_t)a.numel(), "tap %s: buffer has %zu elements, activation %lld",
            name.c_str(), it->second.second, a.numel());
  if (a.f) {
    CUDA_CHECK(cudaMemcpyAsync(it->second.first, a.f, (size_t)a.numel() * sizeof(float), cudaMemcpyDefault, eng.stream));
  } else {   // 16-bit residual stream: widen into a scratch buffer first
    float* tmp = nullptr;
    CUDA_CHECK(cudaMallocAsync(&tmp, (size_t)a.numel() * sizeof(float), eng.stream));
    launch_widen16(a.b, tmp, a.numel(), eng.fp16, eng.stream);
    CUDA_CHECK(cudaMemcpyAsync(it->second.first, tmp, (size_t)a.numel() * sizeof(float), cudaMemcpyDefault, eng.stream));
    CUDA_CHECK(cudaFreeAsync(tmp, eng.stream));
  }
}

Act Model::alloc_act(int n, int h, int w, int c, bool f, bool b) {
  Act a; a.n = n; a.h = h; a.w = w; a.c = c;
  if (f && b && stream16_) f = false;   // block-boundary tensors: the 16-bit copy IS the residual stream
  if (f) a.f = eng.alloc<float>((size_t)a.numel());
  if (b) a.b = eng.alloc<bf16>((size_t)a.numel());
  return a;
}

// Called at the start of every forward graph (UNet step, decode): the dry pass measures how many
// statistic slots the pass needs, the real pass zeroes exactly those with one memset.
void Model::begin_pass() {
  if (!eng.dry && pool_need_)
    CUDA_CHECK(cudaMemsetAsync(pool_, 0, pool_need_, eng.stream));
  pool_off_ = 0;
}

// zero-initialised statistics slot of the current pass (dry pass: only tallies)
void* Model::pool_take(size_t bytes) {
  const size_t off = (pool_off_ + 15) & ~size_t(15);
  pool_off_ = off + bytes;
  if (eng.dry) return reinterpret_cast<void*>(uintptr_t(0x100000) + off);   // never dereferenced
  LDM_CHECK(pool_off_ <= pool_cap_, "statistics pool overflow (%zu of %zu bytes)", pool_off_, pool_cap_);
  return pool_ + off;
}

void Model::gn(const GNW& g, const Act& x, const Act* skip, bool silu, bf16* out) {
  const int cb = skip ? skip->c : 0;
  LDM_CHECK(x.c + cb == g.c, "GroupNorm channel mismatch %d+%d vs %d", x.c, cb, g.c);
  // one-launch cluster kernel: correct but measured ~2 % slower per UNet step than statistics + apply
  // (two cluster barriers and a serial second pass per CTA outweigh the saved launch): opt-in
  static const bool fused = getenv("LDM_B200_GN_FUSED") && getenv("LDM_B200_GN_FUSED")[0] == '1';
  const bool in16 = x.f == nullptr;
  LDM_CHECK(!skip || ((skip->f == nullptr) == in16), "GroupNorm over a concat needs both halves in the same format");
  const void* xa = in16 ? static_cast<const void*>(x.b) : static_cast<const void*>(x.f);
  const void* xb = !skip ? nullptr : (in16 ? static_cast<const void*>(skip->b) : static_cast<const void*>(skip->f));
  if (fused && !in16 && gn_fused_supported(g.c, x.h * x.w, x.n)) {
    eng.launches += 1;
    if (eng.dry) return;
    launch_gn_fused(x.f, x.c, skip ? skip->f : nullptr, cb, x.n, x.h * x.w, g.eps, g.gamma->f32, g.beta->f32,
                    silu ? 1 : 0, out, eng.fp16, eng.stream);
    return;
  }
  static const bool no_pool = getenv("LDM_B200_GN_POOL") && getenv("LDM_B200_GN_POOL")[0] == '0';
  double* st;
  if (no_pool) {
    st = eng.alloc<double>((size_t)x.n * 64);
    if (!eng.dry) CUDA_CHECK(cudaMemsetAsync(st, 0, (size_t)x.n * 64 * sizeof(double), eng.stream));
  } else {
    st = static_cast<double*>(pool_take((size_t)x.n * 64 * sizeof(double)));
  }
  eng.launches += 2;
  if (eng.dry) return;
  launch_gn_stats(xa, x.c, xb, cb, x.n, x.h * x.w, st, eng.stream, in16 ? 1 : 0, eng.fp16, sat_dev_);
  launch_gn_apply(xa, x.c, xb, cb, x.n, x.h * x.w, st, g.eps, g.gamma->f32, g.beta->f32,
                  silu ? 1 : 0, out, eng.fp16, eng.stream, in16 ? 1 : 0);
}

void Model::linear(const bf16* a, long long rows, const LinW& w, const float* bias, int act, const float* residual,
                   float* out_f32, bf16* out_bf16, const bf16* res16) {
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_mat(a, rows, w.k, w.k);
  op.b = view_mat(w.wt, w.n, w.k, w.ld);
  int bk = 0;
  op.add_seg(0, 0, 0, 0, w.k, bk);
  op.W = (int)rows; op.H = 1; op.NB = 1;
  op.N = w.n;
  op.bias = bias; op.act = act; op.residual = residual; op.res16 = res16; op.out_f32 = out_f32; op.out_bf16 = out_bf16;
  op.os_x = w.n; op.os_y = 0; op.os_n = 0;
  eng.gemm(op);
}

// 3x3 SAME conv over a bf16 NHWC activation, 9 shifted TMA boxes (K = 9*cin).
static void conv_segments(GemmOp& op, int cin) {
  int bk = 0;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) op.add_seg(0, ky - 1, kx - 1, 0, cin, bk);
}

Act Model::conv3x3(const Act& x, const LinW& w, const float* bias) {
  Act out = alloc_act(x.n, x.h, x.w, w.n);
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_nhwc(x.b, x.n, x.h, x.w, x.c);
  op.b = view_mat(w.wt, w.n, w.k, w.ld);
  conv_segments(op, x.c);
  op.W = x.w; op.H = x.h; op.NB = x.n;
  op.N = w.n;
  op.bias = bias;
  op.out_f32 = out.f; op.out_bf16 = out.b;
  op.os_x = w.n; op.os_y = (long long)x.w * w.n; op.os_n = (long long)x.h * x.w * w.n;
  eng.gemm(op);
  return out;
}

// Upsample.call: nearest x2 (tf.raw_ops.ResizeNearestNeighbor) then a 3x3 SAME conv (unet.py:44-47,
// autoencoder.py:152-155).  Default: four phase-collapsed 2x2 convs straight over the source image
// (launch_pack_upconv_phase): 4/9 of the FLOPs and no materialised 4x activation.  LDM_B200_UPCONV=materialize
// keeps the literal formulation for A/B runs.
Act Model::upconv(const Act& x, const LinW& w9, const LinW& wp, const float* bias) {
  static const bool materialize = getenv("LDM_B200_UPCONV") && !strcmp(getenv("LDM_B200_UPCONV"), "materialize");
  if (materialize) {
    Act up = alloc_act(x.n, x.h * 2, x.w * 2, x.c, false, true);
    eng.launches++;
    if (!eng.dry) launch_upsample2(x.b, x.n, x.h, x.w, x.c, up.b, eng.stream);
    return conv3x3(up, w9, bias);
  }
  const int c = x.c, cout = wp.n;
  Act out = alloc_act(x.n, 2 * x.h, 2 * x.w, cout);
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_nhwc(x.b, x.n, x.h, x.w, c);
  AView b; b.ptr = wp.wt; b.C = 4 * c; b.W = cout; b.H = 4; b.NB = 1;
  b.sx = 4 * c; b.sy = (long long)cout * 4 * c; b.sn = 4ll * cout * 4 * c;
  op.b = b; op.b_mode = B_PHASE; op.num_phases = 4;
  int bk = 0;
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) op.add_seg(0, dy - 1, dx - 1, 0, c, bk);   // the kernel adds the phase (py, px)
  op.W = x.w; op.H = x.h; op.NB = x.n; op.N = cout;
  op.bias = bias;
  op.out_f32 = out.f; op.out_bf16 = out.b;
  // source pixel (y, x), phase (py, px) -> output pixel (2y + py, 2x + px)
  op.os_x = 2ll * cout; op.os_y = 2ll * (2 * x.w) * cout; op.os_n = 4ll * x.h * x.w * cout;
  op.os_phase_y = 2ll * x.w * cout; op.os_phase_x = cout;
  eng.gemm(op);
  return out;
}

// Downsample: zero pad + 3x3 stride-2 VALID conv.  UNet pads (1,1) on both axes (unet.py:22,26-27): tap (ky,kx) of
// output (oy,ox) reads input (2oy + ky - 1, 2ox + kx - 1); the autoencoder's encoder pads (0,1)
// (autoencoder.py:133): (2oy + ky, 2ox + kx).  No im2col: the A operand is a TMA map with element stride 2, one
// shifted box per tap, out-of-range elements zero-filled by TMA.  LDM_B200_DOWNCONV=im2col keeps the old path.
Act Model::downconv(const Act& x, const LinW& w, const float* bias, int pad_lo) {
  static const bool im2col = getenv("LDM_B200_DOWNCONV") && !strcmp(getenv("LDM_B200_DOWNCONV"), "im2col");
  const int ho = x.h / 2, wo = x.w / 2;
  LDM_CHECK(x.h % 2 == 0 && x.w % 2 == 0, "downsample: odd input size %dx%d", x.h, x.w);
  Act out = alloc_act(x.n, ho, wo, w.n);
  if (im2col && pad_lo == 1) {
    const size_t mk = eng.arena.mark();
    bf16* col = eng.alloc<bf16>((size_t)x.n * ho * wo * 9 * x.c);
    eng.launches++;
    if (!eng.dry) launch_im2col_s2(x.b, x.n, x.h, x.w, x.c, col, eng.stream);
    linear(col, (long long)x.n * ho * wo, w, bias, ACT_NONE, nullptr, out.f, out.b);
    eng.arena.release(mk);
    return out;
  }
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_nhwc(x.b, x.n, x.h, x.w, x.c);
  op.a[0].estride = 2;
  op.b = view_mat(w.wt, w.n, w.k, w.ld);
  int bk = 0;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) op.add_seg(0, ky - pad_lo, kx - pad_lo, 0, x.c, bk);
  op.W = wo; op.H = ho; op.NB = x.n; op.N = w.n;
  op.bias = bias;
  op.out_f32 = out.f; op.out_bf16 = out.b;
  op.os_x = w.n; op.os_y = (long long)wo * w.n; op.os_n = (long long)ho * wo * w.n;
  eng.gemm(op);
  return out;
}

// ResidualBlock.call (unet.py:382-398 / autoencoder.py:43-58).  `skip` != null: the input is the
// virtual concat [x | skip] along C (unet.py:135).
Act Model::resblock(const ResW& r, const Act& x, const Act* skip) {
  const int cin = x.c + (skip ? skip->c : 0);
  LDM_CHECK(cin == r.cin, "resblock: input channels %d vs %d", cin, r.cin);
  Act out = alloc_act(x.n, x.h, x.w, r.cout);
  const size_t mk = eng.arena.mark();
  // GN1 + SiLU -> bf16 (concat materialised here)
  Act a1 = alloc_act(x.n, x.h, x.w, cin, false, true);
  gn(r.gn1, x, skip, true, a1.b);
  // conv1 + bias + time projection (unet.py:384-388)
  Act h1 = alloc_act(x.n, x.h, x.w, r.cout, !stream16_, stream16_);   // feeds GroupNorm 2 only
  {
    GemmOp op;
    op.num_a = 1;
    op.a[0] = view_nhwc(a1.b, x.n, x.h, x.w, cin);
    op.b = view_mat(r.conv1.wt, r.cout, r.conv1.k, r.conv1.ld);
    conv_segments(op, cin);
    op.W = x.w; op.H = x.h; op.NB = x.n; op.N = r.cout;
    op.bias = r.conv1.bias->f32;
    if (r.temb_off >= 0) {
      LDM_CHECK(temb_table_ != nullptr, "resblock: time-embedding table not computed");
      op.bias2 = temb_table_ + r.temb_off;
      op.bias2_stride = tproj_cols_;
      op.bias2_by_img = temb_by_img_ ? 1 : 0;
      op.step_ptr = temb_use_step_ ? step_dev_ : nullptr;
    }
    op.out_f32 = h1.f; op.out_bf16 = h1.b;
    op.os_x = r.cout; op.os_y = (long long)x.w * r.cout; op.os_n = (long long)x.h * x.w * r.cout;
    eng.gemm(op);
  }
  // GN2 + SiLU
  Act a2 = alloc_act(x.n, x.h, x.w, r.cout, false, true);
  gn(r.gn2, h1, nullptr, true, a2.b);
  // conv2 (+ shortcut Dense folded into the K loop) + residual
  {
    GemmOp op;
    op.num_a = 1;
    op.a[0] = view_nhwc(a2.b, x.n, x.h, x.w, r.cout);
    op.b = view_mat(r.conv2.wt, r.cout, r.conv2.k, r.conv2.ld);
    conv_segments(op, r.cout);
    int bk = 9 * r.cout;
    if (r.shortcut) {
      op.a[op.num_a] = view_nhwc(x.b, x.n, x.h, x.w, x.c);
      op.add_seg(op.num_a, 0, 0, 0, x.c, bk);
      op.num_a++;
      if (skip) {
        op.a[op.num_a] = view_nhwc(skip->b, x.n, x.h, x.w, skip->c);
        op.add_seg(op.num_a, 0, 0, 0, skip->c, bk);
        op.num_a++;
      }
      op.bias2 = r.sc_bias->f32;
      op.bias2_stride = 0;
    } else if (x.f) {
      op.residual = x.f;
    } else {
      op.res16 = x.b;   // 16-bit residual stream
    }
    op.W = x.w; op.H = x.h; op.NB = x.n; op.N = r.cout;
    op.bias = r.conv2.bias->f32;
    op.out_f32 = out.f; op.out_bf16 = out.b;
    op.os_x = r.cout; op.os_y = (long long)x.w * r.cout; op.os_n = (long long)x.h * x.w * r.cout;
    eng.gemm(op);
  }
  eng.arena.release(mk);
  return out;
}

// softmax(q k^T * scale) v for all (image, head) pairs (unet.py:280-287): two batched tcgen05
// GEMMs around a row softmax.  q [n,t,heads,d] (row stride q_ld), k [n,tk,heads,d] (row stride
// k_ld, image stride k_sn), vt [n,heads,d,tpad], o [n,t,heads*d] (row stride o_ld).
void Model::attention_core(const bf16* q, long long q_ld, const bf16* k, long long k_ld, long long k_sn, int tk,
                           const bf16* vt, int tpad, int n, int t, int heads, int d, float scale, bf16* o,
                           long long o_ld) {
  if (Engine::attention_supported(d) && !force_unfused_attention) {
    AttnOp op;
    op.q = q; op.q_ld = q_ld; op.k = k; op.k_ld = k_ld; op.k_sn = k_sn; op.vt = vt; op.tpad = tpad;
    op.n = n; op.t = t; op.tk = tk; op.heads = heads; op.d = d; op.scale = scale; op.o = o; op.o_ld = o_ld;
    eng.attention(op);
    return;
  }
  // head dims beyond the fused kernel (autoencoder AttentionBlock, d = 512): two batched tcgen05
  // GEMMs around a row softmax
  const size_t mk = eng.arena.mark();
  float* S = eng.alloc<float>((size_t)n * heads * t * tpad);
  bf16* P = eng.alloc<bf16>((size_t)n * heads * t * tpad);
  {
    GemmOp op;
    op.num_a = 1;
    AView a; a.ptr = q; a.C = d; a.W = t; a.H = heads; a.NB = n; a.sx = q_ld; a.sy = d; a.sn = (long long)t * q_ld;
    a.swap_xy = true;
    AView b; b.ptr = k; b.C = d; b.W = tk; b.H = heads; b.NB = n; b.sx = k_ld; b.sy = d; b.sn = k_sn;
    b.swap_xy = true;
    op.a[0] = a; op.b = b; op.b_mode = B_BATCH;
    int bk = 0;
    op.add_seg(0, 0, 0, 0, d, bk);
    op.W = t; op.H = heads; op.NB = n; op.w_b = GEMM_BM; op.h_b = 1; op.n_b = 1;
    op.N = tk;
    op.out_f32 = S;
    op.os_n = (long long)heads * t * tpad; op.os_y = (long long)t * tpad; op.os_x = tpad;
    eng.gemm(op);
  }
  eng.launches++;
  if (!eng.dry) launch_softmax(S, P, (long long)n * heads * t, tk, tpad, scale, eng.fp16, eng.stream);
  {
    GemmOp op;
    op.num_a = 1;
    AView a; a.ptr = P; a.C = tpad; a.W = t; a.H = heads; a.NB = n; a.sx = tpad; a.sy = (long long)t * tpad;
    a.sn = (long long)heads * t * tpad;
    AView b; b.ptr = vt; b.C = tpad; b.W = d; b.H = heads; b.NB = n; b.sx = tpad; b.sy = (long long)d * tpad;
    b.sn = (long long)heads * d * tpad;
    op.a[0] = a; op.b = b; op.b_mode = B_BATCH;
    int bk = 0;
    op.add_seg(0, 0, 0, 0, tk, bk);
    op.W = t; op.H = heads; op.NB = n; op.w_b = GEMM_BM; op.h_b = 1; op.n_b = 1;
    op.N = d;
    op.out_bf16 = o;
    op.os_n = (long long)t * o_ld; op.os_y = d; op.os_x = o_ld;
    eng.gemm(op);
  }
  eng.arena.release(mk);
}

// Fused q|k|v projection: q,k row-major into qk [rows, 2*inner]; v transposed into vt
// [n, heads, d, tpad] so that P.V reads a K-major B operand.  ln_stats != null: z holds RAW rows and the
// LayerNorm in front of the projection is folded into it (w.ln_cs / w.ln_bias, see apply_folds).
static void qkv_projection(Engine& eng, const bf16* z, int n, int t, int c, const LinW& w, const float* bias, int inner,
                           bf16* qk, bf16* vt, int tpad, const long long* ln_stats = nullptr) {
  GemmOp op;
  op.num_a = 1;
  AView a; a.ptr = z; a.C = c; a.W = t; a.H = 1; a.NB = n; a.sx = c; a.sy = (long long)t * c; a.sn = (long long)t * c;
  op.a[0] = a;
  op.b = view_mat(w.wt, w.n, w.k, w.ld);
  int bk = 0;
  op.add_seg(0, 0, 0, 0, c, bk);
  op.W = t; op.H = 1; op.NB = n;
  op.N = 3 * inner;
  op.n_boundary = 2 * inner;
  op.bias = bias;
  if (ln_stats) { op.ln_stats = ln_stats; op.ln_cs = w.ln_cs; op.ln_c = c; op.bias = w.ln_bias; }
  op.out_bf16 = qk;
  op.os_n = (long long)t * 2 * inner; op.os_y = 0; op.os_x = 2 * inner;
  op.out_tr = vt; op.tr_col0 = 2 * inner;
  op.ts_n = (long long)inner * tpad; op.ts_y = 0; op.ts_c = tpad;
  eng.gemm(op);
}

// SpatialTransformer.call (unet.py:356-365) + BasicTransformerBlock (unet.py:308-314).
// Inside the block the token stream y is 16-bit and updated in place by the three residual linears
// (fp32 accumulate, 16-bit residual read in the epilogue); the block's input / output stay fp32.  The
// three LayerNorms never run as kernels: each producer's epilogue accumulates the rows' (sum, sum sq),
// and the consumer GEMM (q|k|v, q, GEGLU) applies mean / rstd in its own epilogue (apply_folds).
Act Model::spatial_transformer(STW& s, const Act& x) {
  const int n = x.n, t = x.h * x.w, c = s.c, heads = cfg.num_heads, d = s.d;
  const long long rows = (long long)n * t;
  LDM_CHECK(x.c == c && heads * d == c, "spatial transformer: channel mismatch");
  LDM_CHECK(s.ctx_k != nullptr && ctx_rows_ == n, "spatial transformer: context not set for %d rows", n);
  LDM_CHECK(rows < (1ll << 30), "spatial transformer: too many token rows");
  Act out = alloc_act(n, x.h, x.w, c);
  const size_t mk = eng.arena.mark();
  bf16* xn = eng.alloc<bf16>((size_t)rows * c);
  gn(s.gn, x, nullptr, false, xn);
  bf16* y = eng.alloc<bf16>((size_t)rows * c);   // the block's token stream
  bf16* o = eng.alloc<bf16>((size_t)rows * c);
  long long* st1 = static_cast<long long*>(pool_take((size_t)rows * 2 * sizeof(long long)));
  long long* st2 = static_cast<long long*>(pool_take((size_t)rows * 2 * sizeof(long long)));
  long long* st3 = static_cast<long long*>(pool_take((size_t)rows * 2 * sizeof(long long)));
  // y = dense1(groupnorm(x)) (unet.py:358-361), row statistics for LayerNorm1
  auto res_linear = [&](const bf16* a, const LinW& w, const bf16* res, bf16* dst, long long* stats) {
    GemmOp op;
    op.num_a = 1;
    op.a[0] = view_mat(a, rows, w.k, w.k);
    op.b = view_mat(w.wt, w.n, w.k, w.ld);
    int bk = 0;
    op.add_seg(0, 0, 0, 0, w.k, bk);
    op.W = (int)rows; op.H = 1; op.NB = 1;
    op.N = w.n;
    op.bias = w.bias->f32;
    op.res16 = res; op.out_bf16 = dst; op.rs_out = stats;
    op.os_x = w.n;
    eng.gemm(op);
  };
  res_linear(xn, s.d1, nullptr, y, st1);
  const float scale = 1.0f / sqrtf((float)d);
  // ---- self attention: y += attn1(LN1(y)) (unet.py:309-310)
  {
    const size_t mk2 = eng.arena.mark();
    const int tpad = round_up(t, 8);
    bf16* qk = eng.alloc<bf16>((size_t)rows * 2 * c);
    bf16* vt = eng.alloc<bf16>((size_t)n * c * tpad);
    if (tpad != t && !eng.dry) CUDA_CHECK(cudaMemsetAsync(vt, 0, (size_t)n * c * tpad * 2, eng.stream));
    qkv_projection(eng, y, n, t, c, s.a1.qkv, nullptr, c, qk, vt, tpad, st1);
    attention_core(qk, 2 * c, qk + c, 2 * c, (long long)t * 2 * c, t, vt, tpad, n, t, heads, d, scale, o, c);
    res_linear(o, s.a1.out, y, y, st2);
    eng.arena.release(mk2);
  }
  // ---- cross attention against the hoisted context K / V^T: y += attn2(LN2(y), ctx) (unet.py:311-312)
  {
    const size_t mk2 = eng.arena.mark();
    bf16* q = eng.alloc<bf16>((size_t)rows * c);
    {
      GemmOp op;
      op.num_a = 1;
      op.a[0] = view_mat(y, rows, c, c);
      op.b = view_mat(s.a2.qkv.wt, s.a2.qkv.n, c, s.a2.qkv.ld);
      int bk = 0;
      op.add_seg(0, 0, 0, 0, c, bk);
      op.W = (int)rows; op.H = 1; op.NB = 1;
      op.N = c;
      op.ln_stats = st2; op.ln_cs = s.a2.qkv.ln_cs; op.ln_c = c; op.bias = s.a2.qkv.ln_bias;
      op.out_bf16 = q;
      op.os_x = c;
      eng.gemm(op);
    }
    const int tk = cfg.max_seq_len, tpad = round_up(tk, 8);
    attention_core(q, c, s.ctx_k, c, (long long)tk * c, tk, s.ctx_vt, tpad, n, t, heads, d, scale, o, c);
    res_linear(o, s.a2.out, y, y, st3);
    eng.arena.release(mk2);
  }
  // ---- GEGLU feed-forward: z = y + ff(geglu(LN3(y))) (unet.py:313, 322-325, 335-338)
  bf16* z = eng.alloc<bf16>((size_t)rows * c);
  {
    const size_t mk2 = eng.arena.mark();
    bf16* g = eng.alloc<bf16>((size_t)rows * 4 * c);
    {
      GemmOp op;
      op.num_a = 1;
      op.a[0] = view_mat(y, rows, c, c);
      op.b = view_mat(s.geglu.wt, 8 * c, c, s.geglu.ld);
      int bk = 0;
      op.add_seg(0, 0, 0, 0, c, bk);
      op.W = (int)rows; op.H = 1; op.NB = 1;
      op.N = 4 * c; op.gemm_n = 8 * c; op.block_n = s.geglu_bn;
      op.act = ACT_GEGLU;
      op.ln_stats = st3; op.ln_cs = s.geglu.ln_cs; op.ln_c = c;
      op.bias = s.geglu_bias_perm;
      op.out_bf16 = g;
      op.os_x = 4 * c;
      eng.gemm(op);
    }
    res_linear(g, s.ff, y, z, nullptr);
    eng.arena.release(mk2);
  }
  // dense2 + the block's fp32 input residual (unet.py:363-364)
  linear(z, rows, s.d2, s.d2.bias->f32, ACT_NONE, x.f, out.f, out.b, x.f ? nullptr : x.b);
  eng.arena.release(mk);
  return out;
}

// =====================================================================================
// Time embedding -> per-ResBlock projection table (unet.py:126-127,386-387)
// rows = one per distinct timestep; table [rows, tproj_cols_]
// =====================================================================================
void Model::compute_temb_table(const int* t_host, int rows, float* table) {
  // fp32 end to end: [cos|sin] -> Dense+SiLU -> Dense -> per-ResBlock Dense(SiLU(.)) for all blocks at once
  const int mc = cfg.model_channels, td = 4 * mc;
  int* t_dev = nullptr;
  float *emb = nullptr, *h1 = nullptr, *temb = nullptr;
  CUDA_CHECK(cudaMalloc(&t_dev, rows * sizeof(int)));
  CUDA_CHECK(cudaMalloc(&emb, (size_t)rows * mc * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&h1, (size_t)rows * td * sizeof(float)));
  CUDA_CHECK(cudaMalloc(&temb, (size_t)rows * td * sizeof(float)));
  CUDA_CHECK(cudaMemcpyAsync(t_dev, t_host, rows * sizeof(int), cudaMemcpyHostToDevice, eng.stream));
  launch_time_embed(t_dev, rows, mc, emb, eng.stream);
  launch_small_dense_f32(emb, time1_w_, time1_b_->f32, rows, mc, td, 0, 1, h1, eng.stream);   // unet.py:72 silu
  launch_small_dense_f32(h1, time2_w_, time2_b_->f32, rows, td, td, 0, 0, temb, eng.stream);   // unet.py:73
  launch_small_dense_f32(temb, tproj_w_, tproj_bias_, rows, td, tproj_cols_, 1, 0, table, eng.stream);  // unet.py:386
  eng.launches += 4;
  eng.sync();
  cudaFree(t_dev); cudaFree(emb); cudaFree(h1); cudaFree(temb);
}

long long Model::saturated() {
  unsigned long long v = 0;
  CUDA_CHECK(cudaSetDevice(eng.device));
  eng.sync();
  CUDA_CHECK(cudaMemcpy(&v, sat_dev_, sizeof v, cudaMemcpyDeviceToHost));
  return (long long)v;
}

void* Model::stage(int slot, size_t bytes) {
  if (stage_cap_[slot] < bytes) {
    if (stage_ptr_[slot]) { eng.sync(); cudaFree(stage_ptr_[slot]); stage_ptr_[slot] = nullptr; stage_cap_[slot] = 0; }
    const size_t cap = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
    CUDA_CHECK(cudaMalloc(&stage_ptr_[slot], cap));
    stage_cap_[slot] = cap;
  }
  return stage_ptr_[slot];
}

void Model::ensure_events() {
  if (ev0_) return;
  CUDA_CHECK(cudaEventCreate(&ev0_));
  CUDA_CHECK(cudaEventCreate(&ev1_));
}

// =====================================================================================
// Context: hoisted K / V^T projections of all SpatialTransformers (unet.py:276-277)
// =====================================================================================
void Model::set_context(const float* ctx, int n) {
  LDM_CHECK(finalized && model_ready_[1], "set_context: unet weights not finalized");
  CUDA_CHECK(cudaSetDevice(eng.device));
  const int tk = cfg.max_seq_len, cd = cfg.context_dim, tpad = round_up(tk, 8);
  const size_t nel = (size_t)n * tk * cd;
  float* cf = static_cast<float*>(stage(ST_A, nel * sizeof(float)));
  bf16* cb = static_cast<bf16*>(stage(ST_B, nel * sizeof(bf16)));
  bool realloc_ctx = false;
  CUDA_CHECK(cudaMemcpyAsync(cf, ctx, nel * sizeof(float), cudaMemcpyDefault, eng.stream));
  if (cfg.precision == 2) {   // fp32 validation mode: the context stays fp32, K / V are projected per evaluation (validate.cu)
    if (ctx_f32_cap_ < nel) { dev_free(ctx_f32_); ctx_f32_ = dev_alloc<float>(nel); ctx_f32_cap_ = nel; }
    CUDA_CHECK(cudaMemcpyAsync(ctx_f32_, cf, nel * sizeof(float), cudaMemcpyDeviceToDevice, eng.stream));
    ctx_rows_ = n;
    eng.sync();
    return;
  }
  launch_f32_to_bf16(cf, cb, (long long)nel, 0, eng.fp16, eng.stream);
  for (STW* s : all_st_) {
    const int c = s->c;
    if (n > ctx_cap_rows_ || !s->ctx_k) {   // grow-only: a smaller batch reuses a prefix of the buffers
      dev_free(s->ctx_k);
      dev_free(s->ctx_vt);
      s->ctx_k = dev_alloc<bf16>((size_t)n * tk * c);
      s->ctx_vt = dev_alloc<bf16>((size_t)n * c * tpad, true);
      realloc_ctx = true;
    }
    GemmOp op;
    op.num_a = 1;
    AView a; a.ptr = cb; a.C = cd; a.W = tk; a.H = 1; a.NB = n; a.sx = cd; a.sy = (long long)tk * cd; a.sn = (long long)tk * cd;
    op.a[0] = a;
    op.b = view_mat(s->a2.kv.wt, 2 * c, cd, s->a2.kv.ld);
    int bk = 0;
    op.add_seg(0, 0, 0, 0, cd, bk);
    op.W = tk; op.H = 1; op.NB = n; op.w_b = GEMM_BM; op.h_b = 1; op.n_b = 1;
    op.N = 2 * c; op.n_boundary = c;
    op.out_bf16 = s->ctx_k;
    op.os_n = (long long)tk * c; op.os_x = c;
    op.out_tr = s->ctx_vt; op.tr_col0 = c;
    op.ts_n = (long long)c * tpad; op.ts_c = tpad;
    eng.gemm(op);
  }
  ctx_rows_ = n;
  if (n > ctx_cap_rows_) ctx_cap_rows_ = n;
  eng.sync();
  // the captured step reads ctx_k / ctx_vt by address: only new buffers invalidate it
  if (realloc_ctx && step_graph_) { cudaGraphExecDestroy(step_graph_); step_graph_ = nullptr; }
}

// =====================================================================================
// UNet.call (unet.py:118-138)
// =====================================================================================
void Model::unet_eps(const float* x, int nsrc, int n, int h, int w, float* eps_out) {
  const int mc = cfg.model_channels;
  if (cfg.precision == 2) {   // fp32 validation mode: nothing of the 16-bit engine runs
    if (!eng.dry) unet_eps_f32(x, nsrc, n, h, w, eps_out);
    return;
  }
  begin_pass();
  Act cur = alloc_act(n, h, w, mc);
  eng.launches++;
  if (!eng.dry) launch_conv_in(x, nsrc, n, h, w, conv_in_k_->f32, conv_in_b_->f32, mc, cur.f, cur.b, eng.fp16, eng.stream);
  tap("conv_in", cur);
  std::vector<Act> hiddens{cur};
  int bi = 0;
  for (auto& blk : in_blocks_) {
    if (blk.kind == 1) {
      cur = downconv(cur, blk.resample, blk.resample.bias->f32, 1);   // unet.py:22,26-27
    } else {
      cur = resblock(blk.res, cur, nullptr);
      if (bi == 0) tap("in0_res", cur);
      if (blk.has_st) cur = spatial_transformer(blk.st, cur);
    }
    tap("in" + std::to_string(bi++), cur);
    hiddens.push_back(cur);
  }
  cur = resblock(mid_res1_, cur, nullptr);
  cur = spatial_transformer(mid_st_, cur);
  cur = resblock(mid_res2_, cur, nullptr);
  tap("mid", cur);
  bi = 0;
  for (auto& blk : out_blocks_) {
    Act skip = hiddens.back();
    hiddens.pop_back();
    cur = resblock(blk.res, cur, &skip);
    if (blk.has_st) cur = spatial_transformer(blk.st, cur);
    if (blk.has_up) {
      cur = upconv(cur, blk.resample, blk.up_phase, blk.resample.bias->f32);   // unet.py:44-47
    }
    tap("out" + std::to_string(bi++), cur);
  }
  // conv_out(silu(groupnorm(x))) (unet.py:137)
  bf16* a = eng.alloc<bf16>((size_t)cur.numel());
  gn(out_gn_, cur, nullptr, true, a);
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_nhwc(a, n, h, w, mc);
  op.b = view_mat(conv_out_.wt, conv_out_.n, conv_out_.k, conv_out_.ld);
  conv_segments(op, mc);
  op.W = w; op.H = h; op.NB = n; op.N = cfg.out_channels;
  op.bias = conv_out_.bias->f32;
  op.out_f32 = eps_out;
  op.os_x = cfg.out_channels; op.os_y = (long long)w * cfg.out_channels; op.os_n = (long long)h * w * cfg.out_channels;
  eng.gemm(op);
}

void Model::unet_forward(const float* x, const int* t_host, int n, int h, int w, float* eps_out) {
  LDM_CHECK(finalized && model_ready_[1], "unet_forward: unet weights not finalized");
  LDM_CHECK(ctx_rows_ == n, "unet_forward: context set for %d rows, input has %d", ctx_rows_, n);
  CUDA_CHECK(cudaSetDevice(eng.device));
  if (fwd_temb_rows_ < n) {
    dev_free(fwd_temb_);
    fwd_temb_ = dev_alloc<float>((size_t)n * tproj_cols_);
    fwd_temb_rows_ = n;
  }
  compute_temb_table(t_host, n, fwd_temb_);
  temb_table_ = fwd_temb_;
  temb_by_img_ = true; temb_use_step_ = false;
  const size_t nel = (size_t)n * h * w * 4;
  float* xd = static_cast<float*>(stage(ST_A, nel * sizeof(float)));
  float* ed = static_cast<float*>(stage(ST_B, (size_t)n * h * w * cfg.out_channels * sizeof(float)));
  CUDA_CHECK(cudaMemcpyAsync(xd, x, nel * sizeof(float), cudaMemcpyDefault, eng.stream));
  // size the arena with a dry pass, then run
  {
    DryPass dry(eng);
    unet_eps(xd, n, n, h, w, ed);
  }
  ensure_arena(eng.arena.peak());
  eng.arena.reset();
  unet_eps(xd, n, n, h, w, ed);
  CUDA_CHECK(cudaMemcpyAsync(eps_out, ed, (size_t)n * h * w * cfg.out_channels * sizeof(float), cudaMemcpyDefault,
                             eng.stream));
  eng.sync();
}

// =====================================================================================
// Sampler (model_runners.py:438-509)
// =====================================================================================
void Model::configure_sampler(int S, const int* ddim_t, const float* coeffs) {
  LDM_CHECK(finalized && model_ready_[1], "configure_sampler: unet weights not finalized");
  CUDA_CHECK(cudaSetDevice(eng.device));
  S_ = S;
  ddim_t_.assign(ddim_t, ddim_t + S);
  invalidate_graph();
  dev_free(coeffs_dev_);
  dev_free(sampler_temb_);
  coeffs_dev_ = dev_alloc<float>((size_t)S * 8);
  CUDA_CHECK(cudaMemcpyAsync(coeffs_dev_, coeffs, (size_t)S * 8 * sizeof(float), cudaMemcpyHostToDevice, eng.stream));
  sampler_temb_ = dev_alloc<float>((size_t)S * tproj_cols_);
  compute_temb_table(ddim_t, S, sampler_temb_);
  sampler_stale_ = false;
}

void Model::ddim_step(const float* xt, const float* eps2, const float* noise, int index, float guidance, int clip,
                      int b, int h, int w, float* xt_out, float* x0_out) {
  LDM_CHECK(S_ > 0 && index >= 0 && index < S_, "ddim_step: index %d outside [0,%d)", index, S_);
  CUDA_CHECK(cudaSetDevice(eng.device));
  const long long nh = (long long)b * h * w * 4;
  float *dx = static_cast<float*>(stage(ST_A, nh * 4)), *de = static_cast<float*>(stage(ST_B, 2 * nh * 4));
  float *dout = static_cast<float*>(stage(ST_C, nh * 4)), *dn = nullptr, *d0 = nullptr;
  CUDA_CHECK(cudaMemcpyAsync(dx, xt, nh * 4, cudaMemcpyDefault, eng.stream));
  CUDA_CHECK(cudaMemcpyAsync(de, eps2, 2 * nh * 4, cudaMemcpyDefault, eng.stream));
  if (noise) {
    dn = static_cast<float*>(stage(ST_D, nh * 4));
    CUDA_CHECK(cudaMemcpyAsync(dn, noise, nh * 4, cudaMemcpyDefault, eng.stream));
  }
  if (x0_out) d0 = static_cast<float*>(stage(ST_E, nh * 4));
  launch_ddim_update(de, dx, dn, 0, coeffs_dev_, nullptr, index, guidance, clip, dout, d0, nh, eng.stream);
  eng.launches++;
  CUDA_CHECK(cudaMemcpyAsync(xt_out, dout, nh * 4, cudaMemcpyDefault, eng.stream));
  if (x0_out) CUDA_CHECK(cudaMemcpyAsync(x0_out, d0, nh * 4, cudaMemcpyDefault, eng.stream));
  eng.sync();
}

void Model::sample(const float* x_init, const float* noise, int b, int h, int w, float guidance, float* latents_out,
                   float* eps_trace, int steps_limit, int use_graph) {
  LDM_CHECK(S_ > 0, "sample: sampler not configured");
  LDM_CHECK(ctx_rows_ == 2 * b, "sample: context has %d rows, need 2*B = %d", ctx_rows_, 2 * b);
  CUDA_CHECK(cudaSetDevice(eng.device));
  const long long nh = (long long)b * h * w * 4;
  if (xt_cap_ < (size_t)nh) {
    dev_free(xt_dev_);
    dev_free(eps_dev_);
    xt_dev_ = dev_alloc<float>(nh);
    eps_dev_ = dev_alloc<float>(2 * nh);
    xt_cap_ = nh;
    if (step_graph_) { cudaGraphExecDestroy(step_graph_); step_graph_ = nullptr; }
  }
  if (noise && noise_cap_ < (size_t)(nh * S_)) {
    dev_free(noise_dev_);
    noise_dev_ = dev_alloc<float>(nh * S_);
    noise_cap_ = nh * S_;
    if (step_graph_) { cudaGraphExecDestroy(step_graph_); step_graph_ = nullptr; }
  }
  temb_table_ = sampler_temb_;
  temb_by_img_ = false; temb_use_step_ = true;
  // arena sizing
  long long per_step = 0;
  {
    DryPass dry(eng);
    unet_eps(xt_dev_, b, 2 * b, h, w, eps_dev_);
    per_step = dry.launches() + 2;
  }
  ensure_arena(eng.arena.peak());

  ensure_events();
  cudaEvent_t e0 = ev0_, e1 = ev1_;
  CUDA_CHECK(cudaEventRecord(e0, eng.stream));
  CUDA_CHECK(cudaMemcpyAsync(xt_dev_, x_init, nh * 4, cudaMemcpyDefault, eng.stream));
  if (noise) CUDA_CHECK(cudaMemcpyAsync(noise_dev_, noise, nh * S_ * 4, cudaMemcpyDefault, eng.stream));
  int start = S_ - 1;
  CUDA_CHECK(cudaMemcpyAsync(step_dev_, &start, sizeof(int), cudaMemcpyHostToDevice, eng.stream));
  const int nsteps = (steps_limit > 0 && steps_limit < S_) ? steps_limit : S_;
  auto one_step = [&]() {
    eng.arena.reset();
    unet_eps(xt_dev_, b, 2 * b, h, w, eps_dev_);
    launch_ddim_update(eps_dev_, xt_dev_, noise ? noise_dev_ : nullptr, nh, coeffs_dev_, step_dev_, 0, guidance, 0,
                       xt_dev_, nullptr, nh, eng.stream);
    launch_step_advance(step_dev_, -1, eng.stream);
    eng.launches += 2;
  };
  const bool graph_ok = use_graph != 0 && cfg.precision != 2;   // the validation path allocates as it goes
  if (graph_ok) {
    if (!step_graph_ || graph_b_ != b || graph_h_ != h || graph_w_ != w || graph_guid_ != guidance ||
        graph_noise_ != (noise != nullptr)) {
      if (step_graph_) { cudaGraphExecDestroy(step_graph_); step_graph_ = nullptr; }
      eng.sync();
      cudaGraph_t g = nullptr;
      const long long l1 = eng.launches, g1 = eng.gemm_launches, a1 = eng.attn_launches;
      {
        CaptureGuard cap(eng.stream);
        one_step();
        g = cap.end();
      }
      eng.launches = l1; eng.gemm_launches = g1; eng.attn_launches = a1;
      const cudaError_t ie = cudaGraphInstantiate(&step_graph_, g, 0);
      cudaGraphDestroy(g);
      if (ie != cudaSuccess) { step_graph_ = nullptr; CUDA_CHECK(ie); }
      graph_b_ = b; graph_h_ = h; graph_w_ = w; graph_guid_ = guidance; graph_noise_ = noise != nullptr;
    }
    for (int i = 0; i < nsteps; ++i) {
      CUDA_CHECK(cudaGraphLaunch(step_graph_, eng.stream));
      if (eps_trace)   // parity hook: the replayed step's UNet output, copied between two replays
        CUDA_CHECK(cudaMemcpyAsync(eps_trace + (long long)i * 2 * nh, eps_dev_, 2 * nh * 4, cudaMemcpyDefault,
                                   eng.stream));
    }
    eng.launches += per_step * nsteps;
  } else {
    for (int i = 0; i < nsteps; ++i) {
      one_step();
      if (eps_trace)
        CUDA_CHECK(cudaMemcpyAsync(eps_trace + (long long)i * 2 * nh, eps_dev_, 2 * nh * 4, cudaMemcpyDefault,
                                   eng.stream));
    }
  }
  if (latents_out) CUDA_CHECK(cudaMemcpyAsync(latents_out, xt_dev_, nh * 4, cudaMemcpyDefault, eng.stream));
  last_sample_b_ = b; last_sample_h_ = h; last_sample_w_ = w;
  CUDA_CHECK(cudaEventRecord(e1, eng.stream));
  eng.sync();
  CUDA_CHECK(cudaEventElapsedTime(&last_loop_ms, e0, e1));
  last_step_ms = last_loop_ms / nsteps;
}

// =====================================================================================
// Text encoder (transformer.py:254-272, 173-182)
// =====================================================================================
void Model::encode_text(const long long* ids, int rows, float* ctx_out) {
  LDM_CHECK(finalized && model_ready_[0], "encode_text: transformer weights not finalized");
  CUDA_CHECK(cudaSetDevice(eng.device));
  const int T = cfg.max_seq_len, D = cfg.text_hidden, H = cfg.text_heads, S = cfg.text_head_dim, inner = H * S;
  // de-duplicate identical sequences (run_ldm_sampler.py:42-45 tiles two distinct rows B times)
  std::vector<int> uniq_of(rows);
  std::vector<const long long*> uniq;
  for (int r = 0; r < rows; ++r) {
    int f = -1;
    for (size_t u = 0; u < uniq.size(); ++u)
      if (!memcmp(uniq[u], ids + (long long)r * T, T * sizeof(long long))) { f = (int)u; break; }
    if (f < 0) { f = (int)uniq.size(); uniq.push_back(ids + (long long)r * T); }
    uniq_of[r] = f;
  }
  const int n = (int)uniq.size();
  std::vector<long long> uids((size_t)n * T);
  for (int u = 0; u < n; ++u) {
    memcpy(&uids[(size_t)u * T], uniq[u], T * sizeof(long long));
    for (int i = 0; i < T; ++i)
      LDM_CHECK(uids[(size_t)u * T + i] >= 0 && uids[(size_t)u * T + i] < cfg.vocab_size, "token id %lld out of range",
                uids[(size_t)u * T + i]);
  }
  const long long R = (long long)n * T;
  const int tpad = round_up(T, 8);
  auto body = [&](long long* ids_dev, float* x) {
    eng.launches++;
    if (!eng.dry) launch_embed(ids_dev, tok_emb_->f32, pos_emb_->f32, (int)R, T, D, x, eng.stream);
    bf16* z = eng.alloc<bf16>((size_t)R * D);
    bf16* qk = eng.alloc<bf16>((size_t)R * 2 * inner);
    bf16* vt = eng.alloc<bf16>((size_t)n * inner * tpad);
    bf16* o = eng.alloc<bf16>((size_t)R * inner);
    bf16* hbuf = eng.alloc<bf16>((size_t)R * cfg.text_filter);
    if (!eng.dry) CUDA_CHECK(cudaMemsetAsync(vt, 0, (size_t)n * inner * tpad * 2, eng.stream));
    const float scale = 1.0f / sqrtf((float)S);
    for (auto& L : text_layers_) {
      eng.launches++;
      if (!eng.dry) launch_layernorm(x, L.ln_mha.gamma->f32, L.ln_mha.beta->f32, (int)R, D, 1e-5f, z, nullptr, eng.fp16, eng.stream);
      qkv_projection(eng, z, n, T, D, L.attn.qkv, nullptr, inner, qk, vt, tpad);
      attention_core(qk, 2 * inner, qk + inner, 2 * inner, (long long)T * 2 * inner, T, vt, tpad, n, T, H, S, scale, o,
                     inner);
      linear(o, R, L.attn.out, L.attn.out.bias->f32, ACT_NONE, x, x, nullptr);
      eng.launches++;
      if (!eng.dry) launch_layernorm(x, L.ln_ffn.gamma->f32, L.ln_ffn.beta->f32, (int)R, D, 1e-5f, z, nullptr, eng.fp16, eng.stream);
      linear(z, R, L.f1, L.f1.bias->f32, ACT_GELU, nullptr, nullptr, hbuf);
      linear(hbuf, R, L.f2, L.f2.bias->f32, ACT_NONE, x, x, nullptr);
    }
  };
  long long* ids_dev = static_cast<long long*>(stage(ST_C, uids.size() * sizeof(long long)));
  float* x = static_cast<float*>(stage(ST_D, (size_t)R * D * sizeof(float)));
  float* y = static_cast<float*>(stage(ST_E, (size_t)R * D * sizeof(float)));
  CUDA_CHECK(cudaMemcpyAsync(ids_dev, uids.data(), uids.size() * sizeof(long long), cudaMemcpyHostToDevice, eng.stream));
  if (cfg.precision == 2) {   // fp32 validation mode (validate.cu)
    launch_embed(ids_dev, tok_emb_->f32, pos_emb_->f32, (int)R, T, D, x, eng.stream);
    eng.launches++;
    encode_text_f32(x, n, y);
  } else {
    {
      DryPass dry(eng);
      body(ids_dev, x);
    }
    ensure_arena(eng.arena.peak());
    eng.arena.reset();
    body(ids_dev, x);
    launch_layernorm(x, text_ln_.gamma->f32, text_ln_.beta->f32, (int)R, D, 1e-5f, nullptr, y, eng.fp16, eng.stream);
    eng.launches++;
  }
  for (int r = 0; r < rows; ++r)
    CUDA_CHECK(cudaMemcpyAsync(ctx_out + (long long)r * T * D, y + (long long)uniq_of[r] * T * D,
                               (size_t)T * D * sizeof(float), cudaMemcpyDefault, eng.stream));
  eng.sync();
}

// =====================================================================================
// Autoencoder decode (autoencoder.py:291-298,361-364,430-436)
// =====================================================================================
Act Model::ae_attention(AEAttnW& a, const Act& x) {
  const int n = x.n, t = x.h * x.w, c = a.c;
  const long long rows = (long long)n * t;
  Act out = alloc_act(n, x.h, x.w, c);
  const size_t mk = eng.arena.mark();
  bf16* xn = eng.alloc<bf16>((size_t)rows * c);
  gn(a.gn, x, nullptr, false, xn);
  const int tpad = round_up(t, 8);
  bf16* qk = eng.alloc<bf16>((size_t)rows * 2 * c);
  bf16* vt = eng.alloc<bf16>((size_t)n * c * tpad);
  bf16* o = eng.alloc<bf16>((size_t)rows * c);
  if (tpad != t && !eng.dry) CUDA_CHECK(cudaMemsetAsync(vt, 0, (size_t)n * c * tpad * 2, eng.stream));
  qkv_projection(eng, xn, n, t, c, a.qkv, a.qkv_bias, c, qk, vt, tpad);
  attention_core(qk, 2 * c, qk + c, 2 * c, (long long)t * 2 * c, t, vt, tpad, n, t, 1, c, 1.0f / sqrtf((float)c), o, c);
  linear(o, rows, a.out, a.out.bias->f32, ACT_NONE, x.f, out.f, out.b, x.f ? nullptr : x.b);
  eng.arena.release(mk);
  return out;
}

void Model::decode_body(const float* z, int b, int h, int w, float div, float* img_dev, long long* idx_dev) {
  const int zc = cfg.latent_channels;
  LDM_CHECK(zc == 4, "decode: latent_channels must be 4");
  if (cfg.precision == 2) {   // fp32 validation mode (validate.cu)
    if (!eng.dry) decode_body_f32(z, b, h, w, div, img_dev, idx_dev);
    return;
  }
  begin_pass();
  const long long rows = (long long)b * h * w;
  const float* zin = z;
  float pq_div = div;
  if (cfg.ae_kind == 1) {
    // VectorQuantizer on z/scale (quantize.py:57-78); the quantized latents feed post_quant_conv
    float* zq = eng.alloc<float>((size_t)rows * 4);
    eng.launches += 3;
    if (!eng.dry) launch_vq_argmin(z, rows, 4, codebook_->f32, cfg.vq_vocab, div, idx_dev, zq, eng.stream);
    zin = zq;
    pq_div = 1.0f;
  }
  float* pq = eng.alloc<float>((size_t)rows * 4);
  eng.launches++;
  if (!eng.dry) launch_dense4(zin, rows, pq_div, pq_k_->f32, pq_b_->f32, pq, eng.stream);
  const int top = cfg.ae_channels * cfg.ae_mult[cfg.ae_num_mult - 1];
  Act cur = alloc_act(b, h, w, top);
  eng.launches++;
  if (!eng.dry) launch_conv_in(pq, b, b, h, w, ae_conv_in_k_->f32, ae_conv_in_b_->f32, top, cur.f, cur.b, eng.fp16, eng.stream);
  cur = resblock(ae_mid1_, cur, nullptr);
  cur = ae_attention(ae_mid_attn_, cur);
  cur = resblock(ae_mid2_, cur, nullptr);
  for (auto& s : ae_up_) {
    if (s.kind == 0) {
      cur = resblock(s.res, cur, nullptr);
      bool want = false;
      if (cfg.ae_kind == 1)
        for (int k = 0; k < cfg.ae_num_attn_res; ++k) want |= (cfg.ae_attn_res[k] == cur.h);
      LDM_CHECK(!want || s.attn, "decode: attention needed at resolution %d but the autoencoder was built for latent %d",
                cur.h, cfg.ae_build_hw);
      if (want) cur = ae_attention(s.at, cur);
    } else {
      cur = upconv(cur, s.up, s.up_phase, s.up.bias->f32);   // autoencoder.py:152-155
    }
  }
  bf16* a = eng.alloc<bf16>((size_t)cur.numel());
  gn(ae_out_gn_, cur, nullptr, true, a);
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_nhwc(a, b, cur.h, cur.w, cur.c);
  op.b = view_mat(ae_conv_out_.wt, 3, ae_conv_out_.k, ae_conv_out_.ld);
  conv_segments(op, cur.c);
  op.W = cur.w; op.H = cur.h; op.NB = b; op.N = 3;
  op.bias = ae_conv_out_.bias->f32;
  op.out_f32 = img_dev;
  op.os_x = 3; op.os_y = (long long)cur.w * 3; op.os_n = (long long)cur.h * cur.w * 3;
  eng.gemm(op);
}

void Model::decode(const float* z, int b, int h, int w, float div, float* img_out, long long* idx_out) {
  LDM_CHECK(finalized && model_ready_[2], "decode: autoencoder weights not finalized");
  CUDA_CHECK(cudaSetDevice(eng.device));
  if (!z) {   // the latents the last sample() call left on the device
    LDM_CHECK(xt_dev_ && last_sample_b_ == b && last_sample_h_ == h && last_sample_w_ == w,
              "decode(z = NULL): no device-resident latents of shape [%d,%d,%d,4] (last sample: [%d,%d,%d,4])", b, h, w,
              last_sample_b_, last_sample_h_, last_sample_w_);
    z = xt_dev_;
  }
  const long long rows = (long long)b * h * w;
  const long long up = 1ll << (cfg.ae_num_mult - 1);   // the decoder doubles the resolution at every level but the last
  const long long img_el = rows * up * up * 3;
  float* zd = static_cast<float*>(stage(ST_A, rows * 4 * sizeof(float)));
  float* imgd = static_cast<float*>(stage(ST_B, img_el * sizeof(float)));
  long long* idxd = cfg.ae_kind == 1 ? static_cast<long long*>(stage(ST_C, rows * sizeof(long long))) : nullptr;
  {
    DryPass dry(eng);
    decode_body(zd, b, h, w, div, imgd, idxd);
  }
  ensure_arena(eng.arena.peak());
  eng.arena.reset();
  ensure_events();
  CUDA_CHECK(cudaEventRecord(ev0_, eng.stream));
  CUDA_CHECK(cudaMemcpyAsync(zd, z, rows * 4 * sizeof(float), cudaMemcpyDefault, eng.stream));
  decode_body(zd, b, h, w, div, imgd, idxd);
  CUDA_CHECK(cudaMemcpyAsync(img_out, imgd, img_el * sizeof(float), cudaMemcpyDefault, eng.stream));
  if (idx_out && idxd) CUDA_CHECK(cudaMemcpyAsync(idx_out, idxd, rows * sizeof(long long), cudaMemcpyDefault, eng.stream));
  CUDA_CHECK(cudaEventRecord(ev1_, eng.stream));
  eng.sync();
  CUDA_CHECK(cudaEventElapsedTime(&last_decode_ms, ev0_, ev1_));
}

// =====================================================================================
// Autoencoder encode side (autoencoder.py:242-249,354-359,421-425) and get_latents (model_runners.py:602-625)
// =====================================================================================
void Model::encode_body(const float* img, int b, int h, int w, float* moments_dev) {
  if (cfg.precision == 2) {   // fp32 validation mode (validate.cu)
    if (!eng.dry) encode_body_f32(img, b, h, w, moments_dev);
    return;
  }
  begin_pass();
  Act cur = alloc_act(b, h, w, cfg.ae_channels);
  eng.launches++;
  if (!eng.dry) launch_conv_in(img, b, b, h, w, enc_conv_in_k_->f32, enc_conv_in_b_->f32, cfg.ae_channels, cur.f, cur.b,
                               eng.fp16, eng.stream, 3);
  for (auto& s : enc_down_) {
    if (s.kind == 0) {
      cur = resblock(s.res, cur, nullptr);
      bool want = false;
      if (cfg.ae_kind == 1)
        for (int k = 0; k < cfg.ae_num_attn_res; ++k) want |= (cfg.ae_attn_res[k] == cur.h);   // autoencoder.py:116
      LDM_CHECK(!want || s.attn, "encode: attention needed at resolution %d but the autoencoder was built for %d-pixel images",
                cur.h, cfg.ae_build_hw << (cfg.ae_num_mult - 1));
      if (want) cur = ae_attention(s.at, cur);
    } else {
      cur = downconv(cur, s.down, s.down.bias->f32, 0);   // tf.pad (0,1) + stride-2 VALID (autoencoder.py:133-135)
    }
  }
  cur = resblock(enc_mid1_, cur, nullptr);
  cur = ae_attention(enc_mid_attn_, cur);
  cur = resblock(enc_mid2_, cur, nullptr);
  bf16* a = eng.alloc<bf16>((size_t)cur.numel());
  gn(enc_out_gn_, cur, nullptr, true, a);
  const long long rows = (long long)b * cur.h * cur.w;
  float* pre = eng.alloc<float>((size_t)rows * enc_z_);
  GemmOp op;
  op.num_a = 1;
  op.a[0] = view_nhwc(a, b, cur.h, cur.w, cur.c);
  op.b = view_mat(enc_conv_out_.wt, enc_z_, enc_conv_out_.k, enc_conv_out_.ld);
  conv_segments(op, cur.c);
  op.W = cur.w; op.H = cur.h; op.NB = b; op.N = enc_z_;
  op.bias = enc_conv_out_.bias->f32;
  op.out_f32 = pre;
  op.os_x = enc_z_; op.os_y = (long long)cur.w * enc_z_; op.os_n = (long long)cur.h * cur.w * enc_z_;
  eng.gemm(op);
  eng.launches++;
  if (!eng.dry) launch_dense_small(pre, rows, enc_z_, quant_k_->f32, quant_b_->f32, moments_dev, eng.stream);
}

void Model::encode_images(const float* images, int b, int h, int w, const float* noise, float scale, float* moments_out,
                          float* latents_out) {
  LDM_CHECK(finalized && model_ready_[3], "encode: autoencoder encoder weights not finalized");
  CUDA_CHECK(cudaSetDevice(eng.device));
  const int f = 1 << (cfg.ae_num_mult - 1);
  LDM_CHECK(h % f == 0 && w % f == 0, "encode: image %dx%d not divisible by %d", h, w, f);
  const long long rows = (long long)b * (h / f) * (w / f);
  const int z = cfg.latent_channels;
  float* imgd = static_cast<float*>(stage(ST_A, (size_t)b * h * w * 3 * sizeof(float)));
  float* mom = static_cast<float*>(stage(ST_B, (size_t)rows * enc_z_ * sizeof(float)));
  float* lat = static_cast<float*>(stage(ST_C, (size_t)rows * z * sizeof(float)));
  float* nz = noise ? static_cast<float*>(stage(ST_D, (size_t)rows * z * sizeof(float))) : nullptr;
  {
    DryPass dry(eng);
    encode_body(imgd, b, h, w, mom);
  }
  ensure_arena(eng.arena.peak());
  eng.arena.reset();
  ensure_events();
  CUDA_CHECK(cudaEventRecord(ev0_, eng.stream));
  CUDA_CHECK(cudaMemcpyAsync(imgd, images, (size_t)b * h * w * 3 * sizeof(float), cudaMemcpyDefault, eng.stream));
  if (nz) CUDA_CHECK(cudaMemcpyAsync(nz, noise, (size_t)rows * z * sizeof(float), cudaMemcpyDefault, eng.stream));
  encode_body(imgd, b, h, w, mom);
  if (moments_out)
    CUDA_CHECK(cudaMemcpyAsync(moments_out, mom, (size_t)rows * enc_z_ * sizeof(float), cudaMemcpyDefault, eng.stream));
  if (latents_out) {
    launch_posterior_sample(mom, cfg.ae_kind == 0 ? nz : nullptr, rows, z, cfg.ae_kind == 0 ? 1 : 0, scale, lat, eng.stream);
    eng.launches++;
    CUDA_CHECK(cudaMemcpyAsync(latents_out, lat, (size_t)rows * z * sizeof(float), cudaMemcpyDefault, eng.stream));
  }
  CUDA_CHECK(cudaEventRecord(ev1_, eng.stream));
  eng.sync();
  CUDA_CHECK(cudaEventElapsedTime(&last_encode_ms, ev0_, ev1_));
}

void Model::vq_argmin(const float* z, long long rows, float div, long long* idx_out, float* zq_out) {
  LDM_CHECK(codebook_ && codebook_->set, "vq_argmin: codebook not set (autoencoder kind must be vq)");
  CUDA_CHECK(cudaSetDevice(eng.device));
  float* zd = static_cast<float*>(stage(ST_A, rows * 4 * sizeof(float)));
  float* zq = static_cast<float*>(stage(ST_B, rows * 4 * sizeof(float)));
  long long* idxd = static_cast<long long*>(stage(ST_C, rows * sizeof(long long)));
  CUDA_CHECK(cudaMemcpyAsync(zd, z, rows * 4 * sizeof(float), cudaMemcpyDefault, eng.stream));
  launch_vq_argmin(zd, rows, 4, codebook_->f32, cfg.vq_vocab, div, idxd, zq, eng.stream);
  eng.launches += 3;
  CUDA_CHECK(cudaMemcpyAsync(idx_out, idxd, rows * sizeof(long long), cudaMemcpyDefault, eng.stream));
  if (zq_out) CUDA_CHECK(cudaMemcpyAsync(zq_out, zq, rows * 4 * sizeof(float), cudaMemcpyDefault, eng.stream));
  eng.sync();
}

void Model::tensor_to_image(const float* img, int n, long long per, unsigned char* out) {
  CUDA_CHECK(cudaSetDevice(eng.device));
  float* d = static_cast<float*>(stage(ST_A, (size_t)n * per * sizeof(float)));
  unsigned char* o = static_cast<unsigned char*>(stage(ST_B, (size_t)n * per));
  CUDA_CHECK(cudaMemcpyAsync(d, img, (size_t)n * per * sizeof(float), cudaMemcpyDefault, eng.stream));
  launch_tensor_to_image(d, n, per, o, eng.stream);
  eng.launches++;
  CUDA_CHECK(cudaMemcpyAsync(out, o, (size_t)n * per, cudaMemcpyDefault, eng.stream));
  eng.sync();
}

}  // namespace ldm
