// SizeGNN / SizeClassifier.forward (src/linker_size.py:45-91, src/linker_size_lightning.py:83-110): the linker-size
// classifier that runs once per batch right before the sampler (generate.py:88-99).
// Included at the end of dl_engine.cu: it is a second specialisation of the fp32 SIMT kernels of the denoiser
// (k_prep -> [k_edge_simt<GCL, ReLU> -> k_node<ReLU>] x n_layers -> k_sz_out):
//   x = positions * fragment_mask ; h = embedding_in(one_hot * fragment_mask)            (all rows, unmasked: + bias)
//   m_ij = relu(W2 relu(W1 [h_i, h_j, |x_i-x_j|^2] + b1) + b2) * (edge_mask_ij != 0 and |x_i-x_j|^2 < 6)
//   h = (h + W4 relu(W3 [h, sum_j m_ij] + b3) + b4) * fragment_mask          (normalization_factor = 1, 'sum')
//   out[b] = mean over the N padded rows of embedding_out(h)
// It is tiny next to the 500-step sampler (one pass over B*N^2 edges), so the fp32 SIMT kernels are the right tool.
// normalization='batch_norm' (eval mode) is an affine map per channel: the host folds it into W3/b3 and W4/b4.

struct dl_sizegnn {
  dl_sizegnn_config cfg{};
  int num_sms = 0;
  bool finalized = false;
  std::map<std::string, std::vector<float>> raw;
  float* wblob = nullptr;
  std::vector<GclW> layers;
  const float *We_t = nullptr, *be = nullptr, *Wo = nullptr, *bo = nullptr, *zeros = nullptr;
  Workspace ws;
  int64_t launches = 0;
};

namespace {

std::vector<ExpectedParam> sz_expected_params(const dl_sizegnn_config& c) {
  std::vector<ExpectedParam> v;
  v.push_back({"embedding_in.weight", (int64_t)H * c.in_node_nf});
  v.push_back({"embedding_in.bias", H});
  char buf[64];
  for (int l = 0; l < c.n_layers; ++l) {
    snprintf(buf, sizeof(buf), "layer%d.", l);
    std::string p(buf);
    v.push_back({p + "edge_mlp.0.weight", (int64_t)H * (2 * H + 1)});
    v.push_back({p + "edge_mlp.0.bias", H});
    v.push_back({p + "edge_mlp.2.weight", (int64_t)H * H});
    v.push_back({p + "edge_mlp.2.bias", H});
    v.push_back({p + "node_mlp.0.weight", (int64_t)H * 2 * H});
    v.push_back({p + "node_mlp.0.bias", H});
    v.push_back({p + "node_mlp.2.weight", (int64_t)H * H});
    v.push_back({p + "node_mlp.2.bias", H});
  }
  v.push_back({"embedding_out.weight", (int64_t)c.out_node_nf * H});
  v.push_back({"embedding_out.bias", c.out_node_nf});
  return v;
}

dl_status sz_ensure_workspace(dl_sizegnn* e, int B, int N) {
  Workspace& ws = e->ws;
  if (ws.B == B && ws.N == N) return DL_OK;
  free_workspace(ws);
  const size_t n = (size_t)B * N;
  dl_status s;
#define WSA(field, cnt) if ((s = dev_alloc(ws, &ws.field, (cnt))) != DL_OK) return s
  WSA(nm, n); WSA(x0, n * 3); WSA(xa, n * 3); WSA(h, n * H); WSA(ABg, n * 2 * H); WSA(ABgmax, n * 2); WSA(agg, n * H);
  WSA(cls, n); WSA(rowidx, n); WSA(colidx, n); WSA(xrowidx, n); WSA(nr, B); WSA(nc, B); WSA(nxr, B); WSA(n_items, 1);
  WSA(xmols, B); WSA(n_xmols, 1); WSA(items, n); WSA(xitems, n); WSA(n_xitems, 1);
#undef WSA
  ws.B = B; ws.N = N;
  return DL_OK;
}

}  // namespace

extern "C" {

dl_status dl_sizegnn_create(const dl_sizegnn_config* cfg, dl_sizegnn** out) {
  if (!cfg || !out) { set_err("null argument"); return DL_ERR_INVALID; }
  if (cfg->hidden_nf != H) { set_err("SizeGNN: hidden_nf must be %d (got %d)", H, cfg->hidden_nf); return DL_ERR_UNSUPPORTED; }
  if (cfg->in_node_nf < 1 || cfg->in_node_nf > MAX_DIN || cfg->n_layers < 1 || cfg->out_node_nf < 1 ||
      cfg->out_node_nf > SZ_MAX_OUT) {
    set_err("SizeGNN: unsupported shape (in_node_nf %d, n_layers %d, out_node_nf %d)", cfg->in_node_nf, cfg->n_layers,
            cfg->out_node_nf);
    return DL_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_err("no CUDA device: difflinker_b200 has no CPU fallback");
    return DL_ERR_CUDA;
  }
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop{};
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  dl_sizegnn* e = new dl_sizegnn();
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  CK(cudaFuncSetAttribute(k_node<ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * NODE_TM * LDX * sizeof(float)));
  CK(cudaFuncSetAttribute(k_edge_simt<false, ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EDGE_SIMT_SMEM));
  *out = e;
  return DL_OK;
}

dl_status dl_sizegnn_destroy(dl_sizegnn* e) {
  if (!e) return DL_OK;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  free_workspace(e->ws);
  if (e->wblob) cudaFree(e->wblob);
  delete e;
  return DL_OK;
}

dl_status dl_sizegnn_set_weight(dl_sizegnn* e, const char* name, const float* data, int64_t numel) {
  if (!e || !name || !data) { set_err("null argument"); return DL_ERR_INVALID; }
  for (auto& p : sz_expected_params(e->cfg)) {
    if (p.name == name) {
      if (p.numel != numel) {
        set_err("weight %s: expected %lld elements, got %lld", name, (long long)p.numel, (long long)numel);
        return DL_ERR_WEIGHTS;
      }
      e->raw[p.name].assign(data, data + numel);
      e->finalized = false;
      return DL_OK;
    }
  }
  set_err("unexpected weight name %s", name);
  return DL_ERR_WEIGHTS;
}

dl_status dl_sizegnn_finalize_weights(dl_sizegnn* e) {
  if (!e) { set_err("null engine"); return DL_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device));
  for (auto& p : sz_expected_params(e->cfg))
    if (!e->raw.count(p.name)) { set_err("missing weight %s", p.name.c_str()); return DL_ERR_WEIGHTS; }
  const int L = e->cfg.n_layers, F_in = e->cfg.in_node_nf, IN1 = 2 * H + 1;
  Packer pk;
  auto R = [&](const std::string& k) -> const std::vector<float>& { return e->raw[k]; };
  struct Off { size_t W1a, W1b, b1, wd, W2, b2, W3, b3, W4, b4; };
  std::vector<Off> off(L);
  const size_t oWe = pk.add(transpose_block(R("embedding_in.weight"), H, F_in, 0, F_in));
  const size_t obe = pk.add(R("embedding_in.bias"));
  const size_t oWo = pk.add(R("embedding_out.weight"));
  const size_t obo = pk.add(R("embedding_out.bias"));
  const size_t oz = pk.add(std::vector<float>(H, 0.f));
  char buf[64];
  for (int l = 0; l < L; ++l) {
    snprintf(buf, sizeof(buf), "layer%d.", l);
    std::string p(buf);
    const auto& W1 = R(p + "edge_mlp.0.weight");
    Off& o = off[l];
    o.W1a = pk.add(transpose_block(W1, H, IN1, 0, H));
    o.W1b = pk.add(transpose_block(W1, H, IN1, H, H));
    o.b1 = pk.add(R(p + "edge_mlp.0.bias"));
    o.wd = pk.add(column(W1, H, IN1, 2 * H));
    o.W2 = pk.add(transpose_block(R(p + "edge_mlp.2.weight"), H, H, 0, H));
    o.b2 = pk.add(R(p + "edge_mlp.2.bias"));
    o.W3 = pk.add(transpose_block(R(p + "node_mlp.0.weight"), H, 2 * H, 0, 2 * H));
    o.b3 = pk.add(R(p + "node_mlp.0.bias"));
    o.W4 = pk.add(transpose_block(R(p + "node_mlp.2.weight"), H, H, 0, H));
    o.b4 = pk.add(R(p + "node_mlp.2.bias"));
  }
  if (e->wblob) { cudaFree(e->wblob); e->wblob = nullptr; }
  CK(cudaMalloc((void**)&e->wblob, pk.blob.size() * sizeof(float)));
  CK(cudaMemcpy(e->wblob, pk.blob.data(), pk.blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float* base = e->wblob;
  e->We_t = base + oWe; e->be = base + obe; e->Wo = base + oWo; e->bo = base + obo; e->zeros = base + oz;
  e->layers.assign(L, GclW{});
  for (int l = 0; l < L; ++l) {
    const Off& o = off[l];
    GclW w{};
    w.W1a_t = base + o.W1a; w.W1b_t = base + o.W1b; w.b1 = base + o.b1; w.wd = base + o.wd; w.w0 = e->zeros;
    w.W2_t = base + o.W2; w.b2 = base + o.b2; w.W3_t = base + o.W3; w.b3 = base + o.b3; w.W4_t = base + o.W4; w.b4 = base + o.b4;
    e->layers[l] = w;
  }
  e->finalized = true;
  return DL_OK;
}

dl_status dl_sizegnn_forward(dl_sizegnn* e, int32_t B, int32_t N, const float* xh, const int8_t* fragment_mask,
                             const int8_t* edge_mask, float* out, void* stream) {
  if (!e || !e->finalized) { set_err("SizeGNN engine not finalized"); return DL_ERR_INVALID; }
  if (B <= 0 || N <= 0 || !xh || !fragment_mask || !out) { set_err("dl_sizegnn_forward: bad argument"); return DL_ERR_INVALID; }
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dl_status s = sz_ensure_workspace(e, B, N);
  if (s != DL_OK) return s;
  Workspace& ws = e->ws;
  const int n = B * N, L = e->cfg.n_layers;
  Geom gm{};
  gm.B = B; gm.N = N; gm.F = e->cfg.in_node_nf; gm.C = 0; gm.D = e->cfg.in_node_nf;
  gm.graph_type = 4; gm.norm_constant = 0.f; gm.normalization_factor = 1.f;

  CK(cudaMemsetAsync(ws.agg, 0, (size_t)n * H * sizeof(float), st));     // rows without a live edge aggregate to exactly 0
  k_plan_mol<<<B, 256, 2 * N * sizeof(int), st>>>(N, gm.graph_type, edge_mask, fragment_mask, nullptr, ws.rowidx, ws.colidx,
                                                  ws.xrowidx, ws.nr, ws.nc, ws.nxr);
  LAUNCH_CHECK();
  k_plan_items<<<1, 1, 0, st>>>(B, ET, MAXR, 1, MAXR, ws.nr, ws.nc, ws.nxr, ws.items, ws.n_items, ws.xmols, ws.n_xmols, ws.xitems,
                                ws.n_xitems);
  LAUNCH_CHECK();

  const int node_blocks = (n + NODE_TM - 1) / NODE_TM;
  PrepArgs pa{};
  pa.xh = xh; pa.node_mask = fragment_mask; pa.linker_mask = nullptr; pa.t = nullptr; pa.t_numel = 0; pa.context = nullptr;
  pa.We_t = e->We_t; pa.be = e->be;
  pa.proj = ProjW{e->layers[0].W1a_t, e->layers[0].W1b_t, e->layers[0].b1};
  pa.nm = ws.nm; pa.x0 = ws.x0; pa.x = ws.xa; pa.x04 = nullptr; pa.x4 = nullptr; pa.cls = ws.cls; pa.h = ws.h;
  pa.AB = ws.ABg; pa.ABmax = ws.ABgmax;
  k_prep<<<node_blocks, 256, 0, st>>>(gm, pa);
  LAUNCH_CHECK();

  Plan plan{};
  plan.rowidx = ws.rowidx; plan.colidx = ws.colidx; plan.xrowidx = ws.xrowidx; plan.nr = ws.nr; plan.nc = ws.nc; plan.nxr = ws.nxr;
  plan.items = ws.items; plan.n_items = ws.n_items; plan.xmols = ws.xmols; plan.n_xmols = ws.n_xmols;
  plan.xitems = ws.xitems; plan.n_xitems = ws.n_xitems;
  const size_t node_smem = 3 * NODE_TM * LDX * sizeof(float);
  for (int l = 0; l < L; ++l) {
    const GclW& w = e->layers[l];
    EdgeArgs ea{};
    ea.AB = ws.ABg; ea.ABmax = ws.ABgmax; ea.x = ws.xa; ea.x0 = ws.x0; ea.edge_mask = edge_mask; ea.cls = ws.cls; ea.nm = ws.nm;
    ea.W2_t = w.W2_t; ea.b2 = w.b2; ea.wd = w.wd; ea.w0 = w.w0; ea.plan = plan; ea.agg = ws.agg;
    k_edge_simt<false, ACT_RELU><<<e->num_sms, 256, EDGE_SIMT_SMEM, st>>>(gm, ea);
    LAUNCH_CHECK();
    NodeArgs na{};
    na.h = ws.h; na.agg = ws.agg; na.nm = ws.nm; na.W3_t = w.W3_t; na.b3 = w.b3; na.W4_t = w.W4_t; na.b4 = w.b4;
    if (l + 1 < L) {
      const GclW& nx = e->layers[l + 1];
      na.proj1 = ProjW{nx.W1a_t, nx.W1b_t, nx.b1}; na.AB1 = ws.ABg; na.ABmax1 = ws.ABgmax;
    }
    k_node<ACT_RELU><<<node_blocks, 256, node_smem, st>>>(n, na);
    LAUNCH_CHECK();
  }
  k_sz_out<<<B, 256, 0, st>>>(N, e->cfg.out_node_nf, ws.h, e->Wo, e->bo, out);
  LAUNCH_CHECK();
  e->launches += 4 + 2 * L;
  return DL_OK;
}

}  // extern "C"
