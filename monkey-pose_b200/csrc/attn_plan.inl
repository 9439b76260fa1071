// Attention (centre-of-mass) CNN plan: reference attn_model_struct (train_cnn_networks_hgru.py:422-525).
// Included by hgru_lib.cu inside its anonymous namespace (shares fail(), DevBuf, nblk, round_up).
//
// Layer i = conv SxS (S = 3,3,3,3,5) + bias + relu -> max-pool 2x2 -> batch-norm, channels 1 -> w0 -> ... -> w4,
// spatial 128 -> 64 -> 32 -> 16 -> 8 -> 4; then fc (16*w4 -> F) + relu + batch-norm, fc (F -> O).
// Layer 1 is the fused SIMT stem kernel; layers 2..5 and fc_1 are GEMMs on gemm_tc_splitk_kernel over
// im2col operands (bf16 hi/lo, fp32-class accuracy).

struct attn_plan_s {
  int N = 0, H = 0, W = 0, F = 0, O = 0;
  int w[5] = {0, 0, 0, 0, 0};
  bool params_set = false;
  int launches = 0;
  // parameters
  DevBuf w1, b1;                 // layer 1 filters [3][3][1][w0] + bias (fp32, SIMT kernel)
  DevBuf wt[5];                  // [1..4]: conv filters as K-major bf16 hi|lo [Cout][2*Kpad]; wt[0] unused
  DevBuf cb[5];                  // conv biases
  DevBuf fc1_wt, fc1_b, fc2_w, fc2_b;
  DevBuf bn;                     // 6 x (scale, shift), stride bnw
  int bnw = 0;
  // geometry of the GEMM layers (index 1..4 = conv 2..5, index 5 = fc_1)
  int gK[6] = {0}, gKpad[6] = {0}, gM[6] = {0}, gN[6] = {0}, gSplits[6] = {0}, gKbps[6] = {0};
  int gBN[6] = {0};              // N tile of the layer's GEMM: 128 when the output has <= 128 columns, else 256
  CUtensorMap mapA[6], mapB[6];
  // implicit-GEMM convolution (input channels % 64 == 0): 4-D TMA boxes over bf16 hi / lo NHWC activations
  bool implicit_[6] = {false, false, false, false, false, false};
  int gBY[6] = {0}, gNF[6] = {0};
  CUtensorMap mapAhi[6], mapAlo[6];
  DevBuf act_hi[4], act_lo[4];   // bf16 copies of pool[0..3]
  // workspace
  DevBuf resized, pool[5], a_op, part, fc1, out;
  float* bn_scale(int i) { return bn.as<float>() + static_cast<size_t>(2 * i) * bnw; }
  float* bn_shift(int i) { return bn.as<float>() + static_cast<size_t>(2 * i + 1) * bnw; }
  size_t workspace_bytes() const {
    size_t s = w1.bytes + b1.bytes + fc1_wt.bytes + fc1_b.bytes + fc2_w.bytes + fc2_b.bytes + bn.bytes +
               resized.bytes + a_op.bytes + part.bytes + fc1.bytes + out.bytes;
    for (int i = 0; i < 5; ++i) s += wt[i].bytes + cb[i].bytes + pool[i].bytes;
    for (int i = 0; i < 4; ++i) s += act_hi[i].bytes + act_lo[i].bytes;
    return s;
  }
};

static const int kAttnS[5] = {3, 3, 3, 3, 5};        // filter sizes (:443, :456, :469, :482, :495)
static const int kAttnHW = 128;                      // resize target (:442)

static void attn_plan_free(attn_plan_s* p) {
  DevBuf* all[] = {&p->w1, &p->b1, &p->fc1_wt, &p->fc1_b, &p->fc2_w, &p->fc2_b, &p->bn, &p->resized, &p->a_op,
                   &p->part, &p->fc1, &p->out};
  for (auto b : all) b->release();
  for (int i = 0; i < 5; ++i) { p->wt[i].release(); p->cb[i].release(); p->pool[i].release(); }
  for (int i = 0; i < 4; ++i) { p->act_hi[i].release(); p->act_lo[i].release(); }
}

static int attn_plan_build(attn_plan_s* p) {
  int rc = 0;
  auto A = [&](DevBuf& b, size_t bytes) { if (!rc) rc = b.alloc(bytes); };
  const int N = p->N;
  int sms = 0, dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  A(p->w1, sizeof(float) * 9 * p->w[0]); A(p->b1, sizeof(float) * p->w[0]);
  A(p->resized, sizeof(float) * N * kAttnHW * kAttnHW);
  size_t a_max = 0, part_max = 0;
  int hw = kAttnHW / 2;                               // spatial size of layer i's input (i >= 1)
  p->bnw = p->F;
  for (int i = 0; i < 5; ++i) {
    if (p->w[i] > p->bnw) p->bnw = p->w[i];
    A(p->cb[i], sizeof(float) * p->w[i]);
    A(p->pool[i], sizeof(float) * N * (kAttnHW >> (i + 1)) * (kAttnHW >> (i + 1)) * p->w[i]);
  }
  for (int g = 1; g <= 5; ++g) {                      // GEMM layers: conv 2..5, fc_1
    const bool fc = g == 5;
    const int cin = fc ? 16 * p->w[4] : p->w[g - 1];
    const int S = fc ? 1 : kAttnS[g];
    p->gK[g] = S * S * cin;
    p->gKpad[g] = round_up(p->gK[g], hgru::kGemmBK);
    p->gM[g] = fc ? N : N * hw * hw;
    p->gN[g] = fc ? p->F : p->w[g];
    const int total_kb = p->gKpad[g] / hgru::kGemmBK;
    p->gBN[g] = p->gN[g] <= 128 ? 128 : 256;
    const int ntn = (p->gN[g] + p->gBN[g] - 1) / p->gBN[g], ntm = (p->gM[g] + hgru::kGemmBM - 1) / hgru::kGemmBM;
    int splits = sms / (2 * ntn * ntm);               // split K only when the tile grid leaves SMs idle
    if (splits < 1) splits = 1;
    if (splits > total_kb) splits = total_kb;
    p->gKbps[g] = (total_kb + splits - 1) / splits;
    p->gSplits[g] = (total_kb + p->gKbps[g] - 1) / p->gKbps[g];
    const size_t pitch = 2 * static_cast<size_t>(p->gKpad[g]);
    // implicit GEMM when a k-block never straddles a tap (Cin % 64 == 0) and 128-pixel tiles are whole image rows
    const int ppf = hw * hw;
    p->implicit_[g] = !fc && cin % hgru::kGemmBK == 0 && hw <= 64 && hgru::kGemmBM % hw == 0 &&
                      (ppf % hgru::kGemmBM == 0 || hgru::kGemmBM % ppf == 0);
    if (p->implicit_[g]) {
      p->gBY[g] = ppf >= hgru::kGemmBM ? hgru::kGemmBM / hw : hw;
      p->gNF[g] = ppf >= hgru::kGemmBM ? 1 : hgru::kGemmBM / ppf;
      const size_t ab = sizeof(__nv_bfloat16) * static_cast<size_t>(N) * ppf * cin;
      A(p->act_hi[g - 1], ab); A(p->act_lo[g - 1], ab);
    } else {
      a_max = std::max(a_max, sizeof(__nv_bfloat16) * pitch * p->gM[g]);
    }
    part_max = std::max(part_max, sizeof(float) * p->gSplits[g] * static_cast<size_t>(p->gM[g]) * p->gN[g]);
    DevBuf& wbuf = fc ? p->fc1_wt : p->wt[g];
    A(wbuf, sizeof(__nv_bfloat16) * pitch * p->gN[g]);
    if (!rc) CUDA_TRY(cudaMemset(wbuf.p, 0, wbuf.bytes));                 // pad columns stay zero
    if (!fc) hw /= 2;
  }
  A(p->a_op, a_max);
  A(p->part, part_max);
  if (rc) return rc;
  // pad columns [K, Kpad) of the shared A buffer: zero for every layer's pitch (cleared before each im2col
  // only where a layer has padding; done at forward time by a memset of the pad strip -- see attn_forward_impl)
  for (int g = 1; g <= 5; ++g) {
    const size_t pitch = 2 * static_cast<size_t>(p->gKpad[g]);
    DevBuf& wbuf = g == 5 ? p->fc1_wt : p->wt[g];
    int bad = hgru::make_kmajor_bf16_map(&p->mapB[g], wbuf.p, p->gN[g], pitch, p->gBN[g]);
    if (p->implicit_[g]) {
      const int ih = kAttnHW >> g, cin = p->w[g - 1];
      bad |= hgru::make_nhwc_bf16_map(&p->mapAhi[g], p->act_hi[g - 1].p, N, ih, ih, cin, ih, p->gBY[g], p->gNF[g]);
      bad |= hgru::make_nhwc_bf16_map(&p->mapAlo[g], p->act_lo[g - 1].p, N, ih, ih, cin, ih, p->gBY[g], p->gNF[g]);
    } else {
      bad |= hgru::make_kmajor_bf16_map(&p->mapA[g], p->a_op.p, p->gM[g], pitch, hgru::kGemmBM);
    }
    if (bad) return fail(HGRU_E_CUDA, "cuTensorMapEncodeTiled (attention CNN) failed");
  }
  CUDA_TRY(cudaFuncSetAttribute(hgru::gemm_tc_splitk_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                hgru::GemmCfg<256>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(hgru::gemm_tc_splitk_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                hgru::GemmCfg<128>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(hgru::gemm_tc_splitk_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                hgru::GemmCfg<256>::kSmemBytes));
  CUDA_TRY(cudaFuncSetAttribute(hgru::gemm_tc_splitk_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                hgru::GemmCfg<128>::kSmemBytes));
  A(p->fc1_b, sizeof(float) * p->F);
  A(p->fc2_w, sizeof(float) * p->F * p->O); A(p->fc2_b, sizeof(float) * p->O);
  A(p->bn, sizeof(float) * 6 * 2 * p->bnw);
  A(p->fc1, sizeof(float) * N * p->F); A(p->out, sizeof(float) * N * p->O);
  return rc;
}

static int attn_set_params_impl(attn_plan_s* p, const attn_params_t* q, float eps, cudaStream_t st) {
  if (!q) return fail(HGRU_E_INVALID, "attn_set_params: null parameter struct");
  for (int i = 0; i < 5; ++i)
    if (!q->conv_filters[i] || !q->conv_biases[i]) return fail(HGRU_E_INVALID, "attn_set_params: null conv parameter");
  if (!q->fc_1_weights || !q->fc_1_biases || !q->fc_out_weights || !q->fc_out_biases)
    return fail(HGRU_E_INVALID, "attn_set_params: null fc parameter");
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 4; ++j)
      if (!q->bn[i][j]) return fail(HGRU_E_INVALID, "attn_set_params: null batch-norm parameter");
  CUDA_TRY(cudaMemcpyAsync(p->w1.p, q->conv_filters[0], p->w1.bytes, cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < 5; ++i)
    CUDA_TRY(cudaMemcpyAsync(p->cb[i].p, q->conv_biases[i], sizeof(float) * p->w[i], cudaMemcpyDeviceToDevice, st));
  for (int g = 1; g <= 5; ++g) {
    // HWIO [S][S][Cin][Cout] is already [K][Cout] with K = (dy, dx, ci): transpose to K-major bf16 hi|lo
    const float* src = g == 5 ? q->fc_1_weights : q->conv_filters[g];
    DevBuf& wbuf = g == 5 ? p->fc1_wt : p->wt[g];
    dim3 tg((p->gK[g] + 31) / 32, (p->gN[g] + 31) / 32);
    hgru::transpose_to_bf16_kernel<<<tg, 256, 0, st>>>(src, wbuf.as<__nv_bfloat16>(), p->gK[g], p->gN[g],
                                                      p->gKpad[g], 1);
  }
  CUDA_TRY(cudaMemcpyAsync(p->fc1_b.p, q->fc_1_biases, sizeof(float) * p->F, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(p->fc2_w.p, q->fc_out_weights, sizeof(float) * p->F * p->O, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(p->fc2_b.p, q->fc_out_biases, sizeof(float) * p->O, cudaMemcpyDeviceToDevice, st));
  for (int i = 0; i < 6; ++i) {
    const int c = i < 5 ? p->w[i] : p->F;
    hgru::bn_fold_kernel<<<nblk(c), 256, 0, st>>>(q->bn[i][0], q->bn[i][1], q->bn[i][2], q->bn[i][3], eps,
                                                 p->bn_scale(i), p->bn_shift(i), c);
  }
  CUDA_TRY(cudaGetLastError());
  p->params_set = true;
  return 0;
}

static int attn_forward_impl(attn_plan_s* p, const float* frames, float* out, cudaStream_t st) {
  if (!p->params_set) return fail(HGRU_E_STATE, "attn_forward before attn_set_params");
  if (!frames || !out) return fail(HGRU_E_INVALID, "attn_forward: null pointer");
  const int N = p->N;
  p->launches = 0;
  // tf.image.resize_images(input, [128,128])  (:442)
  hgru::resize_bilinear_kernel<<<nblk(static_cast<size_t>(N) * kAttnHW * kAttnHW), 256, 0, st>>>(
      frames, p->resized.as<float>(), N, p->H, p->W, kAttnHW, kAttnHW,
      static_cast<float>(p->H) / static_cast<float>(kAttnHW), static_cast<float>(p->W) / static_cast<float>(kAttnHW));
  // aconv_1 + relu + apool_1 + batch-norm (:443-454): fused SIMT stem kernel (1 input channel)
  {
    const int C = p->w[0], HW = kAttnHW / 2;
    hgru::stem_conv1_pool_bn_kernel<<<dim3(nblk(static_cast<size_t>(N) * HW * ((HW + hgru::kStemPix - 1) / hgru::kStemPix)), C / 8),
                                      256, 0, st>>>(
        p->resized.as<float>(), p->w1.as<float>(), p->cb[0].as<float>(), p->bn_scale(0), p->bn_shift(0),
        p->pool[0].as<float>(), nullptr, N, HW, HW, C, C, 0);
  }
  p->launches += 2;
  if (p->implicit_[1]) {
    const size_t n8 = static_cast<size_t>(N) * (kAttnHW / 2) * (kAttnHW / 2) * p->w[0] / 8;
    hgru::split_bf16_kernel<<<nblk(n8), 256, 0, st>>>(p->pool[0].as<float>(), p->act_hi[0].as<__nv_bfloat16>(),
                                                     p->act_lo[0].as<__nv_bfloat16>(), n8);
    ++p->launches;
  }
  int hw = kAttnHW / 2;
  for (int g = 1; g <= 5; ++g) {
    const bool fc = g == 5;
    const int S = fc ? 1 : kAttnS[g];
    const int cin = fc ? 16 * p->w[4] : p->w[g - 1];
    const int ih = fc ? 1 : hw;
    hgru::GemmArgs ga{p->gM[g], p->gN[g], p->gK[g], p->gKpad[g], p->gKbps[g], p->part.as<float>(),
                      S, cin, ih, ih, p->gBY[g], p->gNF[g]};
    dim3 grid((p->gN[g] + p->gBN[g] - 1) / p->gBN[g], (p->gM[g] + hgru::kGemmBM - 1) / hgru::kGemmBM,
              p->gSplits[g]);
    if (p->implicit_[g]) {
      // conv as an implicit GEMM: the A tile of a k-block is one TMA box of the bf16 activation, shifted by the tap
      if (p->gBN[g] == 128)
        hgru::gemm_tc_splitk_kernel<128, true><<<grid, 256, hgru::GemmCfg<128>::kSmemBytes, st>>>(
            p->mapAhi[g], p->mapAlo[g], p->mapB[g], ga);
      else
        hgru::gemm_tc_splitk_kernel<256, true><<<grid, 256, hgru::GemmCfg<256>::kSmemBytes, st>>>(
            p->mapAhi[g], p->mapAlo[g], p->mapB[g], ga);
      ++p->launches;
    } else {
      if (p->gKpad[g] != p->gK[g])   // the shared operand buffer's pad columns must read as zero for this pitch
        CUDA_TRY(cudaMemsetAsync(p->a_op.p, 0, sizeof(__nv_bfloat16) * 2 * static_cast<size_t>(p->gKpad[g]) * p->gM[g], st));
      const size_t threads = static_cast<size_t>(p->gM[g]) * S * S * (cin / 8);
      hgru::im2col_split_kernel<<<nblk(threads), 256, 0, st>>>(p->pool[g - 1].as<float>(), p->a_op.as<__nv_bfloat16>(),
                                                              N, ih, ih, cin, S, p->gKpad[g]);
      if (p->gBN[g] == 128)
        hgru::gemm_tc_splitk_kernel<128><<<grid, 256, hgru::GemmCfg<128>::kSmemBytes, st>>>(p->mapA[g], p->mapA[g], p->mapB[g], ga);
      else
        hgru::gemm_tc_splitk_kernel<256><<<grid, 256, hgru::GemmCfg<256>::kSmemBytes, st>>>(p->mapA[g], p->mapA[g], p->mapB[g], ga);
      p->launches += 2;
    }
    if (!fc) {
      // bias + relu (:553-554), max-pool (:540-543), batch-norm
      const size_t pt = static_cast<size_t>(N) * (hw / 2) * (hw / 2) * (p->w[g] / 4);
      const bool nxt = g < 4 && p->implicit_[g + 1];      // the next conv reads bf16 hi / lo copies
      hgru::bias_relu_pool_bn_kernel<<<nblk(pt), 256, 0, st>>>(
          p->part.as<float>(), p->gSplits[g], p->cb[g].as<float>(), p->bn_scale(g), p->bn_shift(g),
          p->pool[g].as<float>(), nxt ? p->act_hi[g].as<__nv_bfloat16>() : nullptr,
          nxt ? p->act_lo[g].as<__nv_bfloat16>() : nullptr, N, hw, hw, p->w[g]);
      ++p->launches;
      hw /= 2;
    }
  }
  // afc_1 bias, relu, batch-norm (R-D5: last axis), afc_out (:501-525)
  hgru::fc_tail_kernel<<<N, 256, sizeof(float) * (p->F + hgru::kFcTailScratch), st>>>(
      p->part.as<float>(), p->gSplits[5], p->fc1_b.as<float>(), p->bn_scale(5), p->bn_shift(5),
      p->fc2_w.as<float>(), p->fc2_b.as<float>(), p->fc1.as<float>(), out, N, p->F, p->O);
  ++p->launches;
  CUDA_TRY(cudaGetLastError());
  return 0;
}
