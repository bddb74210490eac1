/*
 * qnnb200.h -- C ABI of libqnnb200.so: the B200 (sm_100a) implementation of the
 * quantized-layer forward path of victorjoos/QuantizedNeuralNetworks-Keras-Tensorflow.
 *
 * The reference has no FFI layer: its seam is the Keras Layer protocol
 * (layers/quantized_layers.py:32-206, layers/binary_layers.py:31-199,
 * layers/ternary_layers.py:30-186) driven by models/model_factory.py:18-72.  Each entry
 * point below replaces the TensorFlow op sequence one of those `call()` methods expands
 * to; the reference line it stands in for is cited on the declaration.
 *
 * Conventions
 *  - every pointer argument named x / w / y / bias / ... is a DEVICE pointer (HBM);
 *    `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - tensors are channels_last (NHWC), the reference's only supported layout (README.md:17);
 *  - all functions return 0 on success, a negative QNNB_E* code otherwise; the message is
 *    available from qnnb_last_error() (thread-local).  Nothing throws across the ABI;
 *  - an unsupported shape is an error, never a silent CPU fallback: there is no CPU path;
 *  - the compute entry points allocate no device memory and keep no global mutable state:
 *    calls are stream-ordered and re-entrant (only qnnb_peer_alloc, which exists to allocate, does).
 */
#ifndef QNNB200_H_
#define QNNB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QNNB_VERSION 100

/* status codes */
#define QNNB_OK            0
#define QNNB_EINVAL       -1   /* bad argument / unsupported shape */
#define QNNB_ECUDA        -2   /* CUDA runtime / driver error      */
#define QNNB_EUNSUPPORTED -3   /* valid request this build cannot serve (e.g. forced impl) */

/* storage kinds of an activation tensor */
#define QNNB_KIND_NONE -1
#define QNNB_KIND_U8    0      /* uint8 pixel levels 0..255, value = level/255 (utils/load_data.py:40) */
#define QNNB_KIND_I8    1      /* int8 levels k, value = k / 2^(abits-1)  (quantized_ops.py:87-100)  */
#define QNNB_KIND_B1    2      /* bit-packed +-1: uint32 words [..][ceil(C/32)], bit c%32 of word c/32 set <=> +1 */
#define QNNB_KIND_F32   3      /* float32 values */

/* weight quantiser (K0) */
#define QNNB_W_QUANT    0      /* quantize(W, nb)   layers/quantized_ops.py:49-66 */
#define QNNB_W_BINARY   1      /* binarize(W, H)    layers/binary_ops.py:54-64    */
#define QNNB_W_TERNARY  2      /* ternarize(W, H)   layers/ternary_ops.py:15-41   */
#define QNNB_W_FLOAT    3      /* no quantiser: the plain Conv2D / Dense of network_type 'float' (models/model_factory.py:24-27) */

/* packed weight formats */
#define QNNB_WFMT_I8    0      /* int8 levels  [cout][kh][kw][cin_pad], cin_pad = cin rounded up to 4, zero filled */
#define QNNB_WFMT_B1    1      /* uint32 words [cout][kh][kw][ceil(cin/32)], pad bits zero */
#define QNNB_WFMT_F32   2      /* float        [cout][kh][kw][cin_pad] (QNNB_W_FLOAT only; consumed with in_kind F32 and w_f32 = 1) */

/* fused activation */
#define QNNB_ACT_NONE   0      /* fp32 out */
#define QNNB_ACT_QUANT  1      /* quantized_tanh(., abits) -> int8 levels   (quantized_ops.py:87-100) */
#define QNNB_ACT_SIGN   2      /* binary_tanh -> packed bits (+1 <=> z > 2^-24) (binary_ops.py:37-51) */
#define QNNB_ACT_LEAKY  3      /* LeakyReLU(alpha) fp32 out                  (model_factory.py:27,34) */
#define QNNB_ACT_SIGN_I8 4     /* binary_tanh -> int8 levels +1 / -1 (same decision as QNNB_ACT_SIGN): the layout that lets the
                                  next BinaryConv2D run on the int8 tensor cores (binary_ops.py:37-51) */

/* kernel selection (testing / profiling) */
#define QNNB_IMPL_AUTO    0
#define QNNB_IMPL_GENERIC 1    /* CUDA-core dp4a / popc / FFMA tiles      */
#define QNNB_IMPL_TCGEN05 2    /* tcgen05.mma kind::i8 + TMA + TMEM (halo-resident implicit GEMM) */
#define QNNB_IMPL_TCGEN05_V1 3 /* first-generation tcgen05 kernel (one TMA box per filter tap); kept for A/B profiling */

/*
 * The fused epilogue applied to every accumulator (fixed op order, every step a separate
 * round-to-nearest fp32 op, no FMA contraction):
 *   c = float(acc) * acc_scale
 *   p = c + bias[ch]                         (bias != NULL)       K.bias_add, quantized_layers.py:185-189
 *   y = p * bn_inv[ch] + bn_shift[ch]        (bn_inv != NULL)     BatchNormalization, models/vgg.py:16
 *   z = (residual + y) * res_mul             (res_kind != NONE)   add + Lambda(x*0.5), models/resnet.py:127-128
 *   act(z)                                                         models/model_factory.py:19-20,36,47
 *   2x2 max-pool (pool == 2)                                       MaxPooling2D, models/vgg.py:23,30,37
 */
typedef struct qnnb_epilogue {
  float        acc_scale;
  const float* bias;        /* [cout] or NULL */
  const float* bn_inv;      /* [cout] or NULL: gamma / sqrt(var + eps) */
  const float* bn_shift;    /* [cout]        : beta - mean * bn_inv    */
  int32_t      res_kind;    /* QNNB_KIND_NONE | QNNB_KIND_I8 | QNNB_KIND_F32 */
  const void*  residual;    /* NHWC, same shape as the (un-pooled) conv output */
  float        res_scale;   /* I8 residual: value = level * res_scale */
  float        res_mul;
  int32_t      act;         /* QNNB_ACT_* ; decides the output kind */
  int32_t      abits;       /* QNNB_ACT_QUANT */
  float        leaky_alpha; /* QNNB_ACT_LEAKY */
  int32_t      pool;        /* 0 | 2 */
} qnnb_epilogue;

/* 2-D convolution, padding='same' (TensorFlow rule, asymmetric under stride 2). */
typedef struct qnnb_conv_desc {
  int32_t n, h, w, cin;     /* input NHWC */
  int32_t cout, kh, kw;     /* filter     */
  int32_t stride;           /* 1 | 2      */
  int32_t in_kind;          /* QNNB_KIND_U8 | I8 | B1 | F32 */
  int32_t impl;             /* QNNB_IMPL_* */
  qnnb_epilogue epi;
  int32_t w_f32;            /* 1: `w` is a QNNB_WFMT_F32 kernel (fp32 values, 'float' networks); needs in_kind F32.  0: levels */
  int32_t max_ctas;         /* SM share: 0 = the persistent kernels take one CTA per SM of the device; k > 0 = at most k CTAs, so
                               that the kernels of independent batches running on other streams sit NEXT to this one (each on
                               its own SMs) and fill each other's pipeline fill / drain and last partial wave */
} qnnb_conv_desc;

/* Dense layer  y[n][u] = epilogue( sum_f x[n][f] * w[u][f] ), optional softmax. */
typedef struct qnnb_dense_desc {
  int32_t n, fin, units;
  int32_t in_kind;          /* QNNB_KIND_I8 | B1 | F32 */
  int32_t softmax;          /* 1: y = softmax(z) and, if logits != NULL, logits = z */
  qnnb_epilogue epi;        /* act must be QNNB_ACT_NONE, pool 0, no residual */
  int32_t avg_positions;    /* 0 | 1: plain.  P > 1 (fp32 input only): x is [n][P][fin] and the layer sees the SUM over the P
                               positions -- AveragePooling2D(8) + Flatten of models/resnet.py:134-135 folded in; acc_scale
                               carries the 1/P */
  int32_t w_f32;            /* 1: `w` is a QNNB_WFMT_F32 kernel; needs in_kind F32 */
  int32_t max_ctas;         /* SM share, as in qnnb_conv_desc */
} qnnb_dense_desc;

int         qnnb_version(void);
const char* qnnb_last_error(void);
/* SM count and compute capability of the current device. */
int         qnnb_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/*
 * K0 -- quantise + pack a kernel once at load time (the reference re-runs quantize /
 * binarize / ternarize on every forward: quantized_layers.py:80,165; binary_layers.py:79,161;
 * ternary_layers.py:78,157).
 *   w_hwio : fp32 (kh,kw,cin,cout) -- Keras HWIO; a Dense kernel (in,units) is kh=kw=1.
 *   out    : QNNB_WFMT_I8 -> int8  [cout][kh][kw][cin_pad]; QNNB_WFMT_B1 -> uint32 [cout][kh][kw][ceil(cin/32)];
 *            QNNB_WFMT_F32 (mode QNNB_W_FLOAT: the kernel values themselves, re-laid out) -> float [cout][kh][kw][cin_pad]
 *   scratch: >= 2 floats of device memory (ternary cutoff); may be NULL for the other modes.
 */
int qnnb_pack_weights(int32_t mode, int32_t nb, float H, const float* w_hwio,
                      int32_t kh, int32_t kw, int32_t cin, int32_t cout,
                      int32_t wfmt, void* out, float* scratch, void* stream);

/* size in bytes of a packed kernel */
int64_t qnnb_packed_weight_bytes(int32_t wfmt, int32_t kh, int32_t kw, int32_t cin, int32_t cout);

/*
 * QuantizedConv2D.call / BinaryConv2D.call / TernaryConv2D.call
 * (quantized_layers.py:164-194, binary_layers.py:160-187, ternary_layers.py:156-174) with the
 * layers that follow it in models/vgg.py / models/resnet.py fused into the epilogue.
 *   x : in_kind U8/I8 -> bytes [n][h][w][cin]; B1 -> uint32 [n][h][w][ceil(cin/32)]; F32 -> float
 *   w : packed by qnnb_pack_weights (QNNB_WFMT_B1 iff in_kind == B1)
 *   y : act QUANT / SIGN_I8 -> int8 [n][oh][ow][cout]; SIGN -> uint32 [n][oh][ow][ceil(cout/32)];
 *       NONE/LEAKY -> float [n][oh][ow][cout]; (oh,ow) after the optional pool.
 */
int qnnb_conv2d(const qnnb_conv_desc* desc, const void* x, const void* w, void* y, void* stream);

/* 1 when qnnb_conv2d would run this descriptor on the tcgen05 tensor-core kernels (QNNB_IMPL_AUTO), else 0; the
 * host uses it to choose between the bit-packed (QNNB_ACT_SIGN) and the int8 (QNNB_ACT_SIGN_I8) form of a +-1 map. */
int qnnb_conv2d_tc_supported(const qnnb_conv_desc* desc);

/* output spatial size of qnnb_conv2d (after pooling) */
int qnnb_conv2d_out_shape(const qnnb_conv_desc* desc, int32_t* oh, int32_t* ow);

/*
 * QuantizedDense.call / BinaryDense.call / TernaryDense.call (quantized_layers.py:79-88,
 * binary_layers.py:78-85, ternary_layers.py:77-84) + the BatchNormalization of models/vgg.py:42
 * or the softmax of models/resnet.py:137.
 *   x : I8 -> int8 [n][fin]; B1 -> uint32 [n][ceil(fin/32)]; F32 -> float [n][fin]
 *   w : packed as a 1x1 kernel with cin = fin, cout = units
 *   y : float [n][units]; logits: float [n][units] or NULL
 */
int qnnb_dense(const qnnb_dense_desc* desc, const void* x, const void* w, float* y, float* logits, void* stream);

/*
 * Un-fused activation quantisers on fp32 tensors (entry for stand-alone use of the ops):
 *   QNNB_ACT_QUANT: quantized_tanh -> int8 levels   (quantized_ops.py:87-100)
 *   QNNB_ACT_SIGN : binary_tanh    -> packed bits   (binary_ops.py:37-51); `channels` = innermost extent
 */
int qnnb_quantize_act(int32_t act, int32_t abits, const float* x, int64_t count, int32_t channels,
                      void* y, void* stream);

/* Stand-alone fp32 layer ops used when a layer object is called outside a fused plan. */
int qnnb_batchnorm_f32(const float* x, int64_t rows, int32_t ch, const float* inv, const float* shift, float* y, void* stream);
int qnnb_maxpool2_f32(const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* y, void* stream);
int qnnb_leaky_f32(const float* x, int64_t count, float alpha, float* y, void* stream);
/* round_through forward value: tf.round, half-to-even (layers/quantized_ops.py:8-14, layers/binary_ops.py:8-13) */
int qnnb_round_f32(const float* x, int64_t count, float* y, void* stream);
/* int8 levels / packed bits -> fp32 values (level * scale, or +-1) */
int qnnb_dequantize(int32_t kind, const void* x, int64_t count, int32_t channels, float scale, float* y, void* stream);

/*
 * Whole-network launch for the small VGG nets: every layer of models/vgg.py:15-42 -- first conv on the uint8 image,
 * (conv -> BatchNormalization -> Activation [-> MaxPooling2D]) x nconv, Flatten, Fc, BatchNormalization -- in ONE
 * kernel; an image stays in one SM's shared memory from the input bytes to the logits (csrc/net_fused.cu).  Covers
 * nets whose packed kernels and activation maps fit there: images up to 32x32 with 1 or 3 channels, 3x3 stride-1
 * convolutions with 32 or 64 filters, quantized_tanh / binary_tanh activations kept as int8 levels, <= 32 units.
 * Same arithmetic as the per-layer entry points (bit-identical results).  Two steps, like qnnb_pack_weights + qnnb_conv2d:
 * qnnb_vgg_pack builds the net's resident image once per set of weights -- every layer's kernel in tensor-core operand
 * order, the dense kernel, the per-channel epilogue constants, qnnb_vgg_blob_bytes(desc) bytes of device memory -- and
 * qnnb_vgg_forward (same descriptor; its weight / bias / BN pointers are not read again) brings it into each SM with
 * bulk copies and runs the batch.
 *   conv[l].w   : packed by qnnb_pack_weights (QNNB_WFMT_I8), cin = image channels (l = 0) or conv[l-1].cout
 *   conv[l].epi : acc_scale, bias, bn_inv / bn_shift, act (QNNB_ACT_QUANT + abits | QNNB_ACT_SIGN_I8); `pool` of the
 *                 epilogue is ignored in favour of conv[l].pool
 *   dense_w     : packed 1x1 kernel [units][fin], fin = (final map) h * w * cout in Flatten (HWC) order
 *   x : uint8 [n][h][w][cin], 16-byte aligned;  y : float [n][units]
 */
#define QNNB_NET_MAX_CONVS 6
typedef struct qnnb_net_conv {
  int32_t       cout;
  int32_t       pool;       /* 0 | 2 */
  const void*   w;
  qnnb_epilogue epi;
} qnnb_net_conv;

typedef struct qnnb_vgg_desc {
  int32_t       n, h, w, cin;
  int32_t       nconv;
  qnnb_net_conv conv[QNNB_NET_MAX_CONVS];
  int32_t       units;
  const void*   dense_w;
  qnnb_epilogue dense_epi;  /* act NONE */
  int32_t       max_ctas;   /* SM share, as in qnnb_conv_desc (0 = min(n, SMs) CTAs, one image each at a time) */
} qnnb_vgg_desc;

/* 1 when qnnb_vgg_forward covers this net, else 0 (the host then runs the per-layer plan) */
int qnnb_vgg_forward_supported(const qnnb_vgg_desc* desc);
int64_t qnnb_vgg_blob_bytes(const qnnb_vgg_desc* desc);          /* 0 when the net is not covered */
int qnnb_vgg_pack(const qnnb_vgg_desc* desc, void* blob, void* stream);
int qnnb_vgg_forward(const qnnb_vgg_desc* desc, const void* blob, const void* x, float* y, void* stream);

/*
 * NVLink logit path for batch-sharded inference (one process per GPU of one box): the reference evaluates a test set
 * with one model.predict / model.evaluate over the whole array (test_resnet.py:63-69); sharded over g GPUs the only
 * exchange is the [N/g, classes] logit block of each shard.  Instead of a collective per step, the gathering rank
 * exports a device buffer (CUDA IPC) and every other rank maps it and passes `mapped + row offset` as the `y` of its
 * final qnnb_dense call: the logits cross NVLink / NVSwitch as the kernel's own stores.
 *   qnnb_peer_alloc : cudaMalloc `bytes` (zero-filled) on the current device, handle = QNNB_PEER_HANDLE_BYTES opaque bytes
 *   qnnb_peer_open  : map a sibling process's buffer for kernels of the CURRENT device (peer access enabled lazily)
 *   qnnb_peer_close / qnnb_peer_free : undo open / alloc
 */
#define QNNB_PEER_HANDLE_BYTES 64
int qnnb_peer_alloc(int64_t bytes, void** ptr, void* handle);
int qnnb_peer_open(const void* handle, void** ptr);
int qnnb_peer_close(void* ptr);
int qnnb_peer_free(void* ptr);

/*
 * Profiling aid (not part of the inference path): register a device buffer of `nwords` uint64 (word 0 = event
 * counter, must be zeroed) in which CTA 0 of the tcgen05 conv kernel records (tag<<32|tile, globaltimer ns) pairs
 * for its pipeline events; NULL switches tracing off.  Process-wide, not thread-safe.
 */
int qnnb_debug_set_trace(void* buf, int64_t nwords);

#ifdef __cplusplus
}
#endif
#endif /* QNNB200_H_ */
