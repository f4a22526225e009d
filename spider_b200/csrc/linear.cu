// The projections either side of the attention (to_q / to_k / to_v / to_out[0] of diffusers' Attention module,
// StoryDiffusion/Comic_Generation.py:155,164-165,185) and the batch runner that issues a whole processor call —
// projections, K/V gather or peer exchange, attention, output projection — from ONE call into the library.
//
// The GEMMs are plain library GEMMs (cuBLASLt, fp32 accumulation, optional bias epilogue): y = x w^T + b with the
// row-major layouts torch's nn.Linear uses, so the weights are read where the module keeps them.  What this buys is
// host time: a processor call through nn.Linear / ctypes wrappers costs ~180 us of Python and launch overhead, which
// bounds the step once the per-GPU work is small (4-frame story on 4-8 GPUs); the batch costs one ctypes transition.
#include <cublasLt.h>

#include <mutex>
#include <unordered_map>

#include "csa_internal.h"

namespace csa {

struct LinearKey {
  int dev, dtype, bias;
  int64_t m, n, k, ldx, ldw, ldy;
  bool operator==(const LinearKey& o) const {
    return dev == o.dev && dtype == o.dtype && bias == o.bias && m == o.m && n == o.n && k == o.k && ldx == o.ldx &&
           ldw == o.ldw && ldy == o.ldy;
  }
};
struct LinearKeyHash {
  size_t operator()(const LinearKey& k) const {
    size_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
    mix(k.dev); mix(k.dtype); mix(k.bias); mix(k.m); mix(k.n); mix(k.k); mix(k.ldx); mix(k.ldw); mix(k.ldy);
    return h;
  }
};
struct LinearPlan {
  cublasLtMatmulDesc_t op = nullptr;
  cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr;
  cublasLtMatmulAlgo_t algo;
  size_t ws = 0;
};

static std::mutex g_lin_mu;
static cublasLtHandle_t g_lt[64];
static std::unordered_map<LinearKey, LinearPlan, LinearKeyHash> g_plans;

static const char* lt_status(cublasStatus_t s) {
  switch (s) {
    case CUBLAS_STATUS_SUCCESS: return "success";
    case CUBLAS_STATUS_NOT_INITIALIZED: return "not initialized";
    case CUBLAS_STATUS_ALLOC_FAILED: return "alloc failed";
    case CUBLAS_STATUS_INVALID_VALUE: return "invalid value";
    case CUBLAS_STATUS_ARCH_MISMATCH: return "arch mismatch";
    case CUBLAS_STATUS_EXECUTION_FAILED: return "execution failed";
    case CUBLAS_STATUS_INTERNAL_ERROR: return "internal error";
    case CUBLAS_STATUS_NOT_SUPPORTED: return "not supported";
    default: return "cublas error";
  }
}

#define LT_TRY(expr)                                                                                       \
  do {                                                                                                     \
    cublasStatus_t st_ = (expr);                                                                           \
    if (st_ != CUBLAS_STATUS_SUCCESS)                                                                      \
      return set_error(1000 + static_cast<int>(st_), "csa_linear: %s -> %s", #expr, lt_status(st_));      \
  } while (0)

// Row-major y[M,N] = x[M,K] w[N,K]^T is, in cuBLAS's column-major terms, C(N x M) = op(A)(N x K) * B(K x M) with
// A = w seen as a K x N matrix (lda = ldw) transposed, B = x seen as K x M (ldb = ldx), C = y seen as N x M.
static int get_plan(const LinearKey& key, size_t ws_bytes, cublasLtHandle_t lt, const LinearPlan** out) {
  auto it = g_plans.find(key);
  if (it != g_plans.end() && it->second.ws <= ws_bytes) {
    *out = &it->second;
    return 0;
  }
  LinearPlan p;
  const cudaDataType_t dt = key.dtype == CSA_DTYPE_BF16 ? CUDA_R_16BF : CUDA_R_16F;
  LT_TRY(cublasLtMatmulDescCreate(&p.op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
  const cublasOperation_t ta = CUBLAS_OP_T, tb = CUBLAS_OP_N;
  LT_TRY(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)));
  LT_TRY(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)));
  if (key.bias) {
    const cublasLtEpilogue_t epi = CUBLASLT_EPILOGUE_BIAS;
    LT_TRY(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &epi, sizeof(epi)));
    LT_TRY(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &dt, sizeof(dt)));
  }
  LT_TRY(cublasLtMatrixLayoutCreate(&p.a, dt, key.k, key.n, key.ldw));
  LT_TRY(cublasLtMatrixLayoutCreate(&p.b, dt, key.k, key.m, key.ldx));
  LT_TRY(cublasLtMatrixLayoutCreate(&p.c, dt, key.n, key.m, key.ldy));
  cublasLtMatmulPreference_t pref = nullptr;
  LT_TRY(cublasLtMatmulPreferenceCreate(&pref));
  LT_TRY(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_bytes,
                                              sizeof(ws_bytes)));
  cublasLtMatmulHeuristicResult_t res;
  int found = 0;
  cublasStatus_t st = cublasLtMatmulAlgoGetHeuristic(lt, p.op, p.a, p.b, p.c, p.c, pref, 1, &res, &found);
  cublasLtMatmulPreferenceDestroy(pref);
  if (st != CUBLAS_STATUS_SUCCESS || found == 0)
    return set_error(1000 + static_cast<int>(st), "csa_linear: no cuBLASLt algorithm for M=%lld N=%lld K=%lld (%s)",
                     (long long)key.m, (long long)key.n, (long long)key.k, lt_status(st));
  p.algo = res.algo;
  p.ws = res.workspaceSize;
  g_plans[key] = p;   // (a replaced plan's descriptors are leaked: a handful per process)
  *out = &g_plans[key];
  return 0;
}

}  // namespace csa

using namespace csa;

extern "C" int csa_linear(const csa_linear_args_t* a, void* stream) {
  if (!a) return set_error(CSA_E_BADARG, "csa_linear: null args");
  if (a->struct_size != sizeof(csa_linear_args_t))
    return set_error(CSA_E_BADARG, "csa_linear: struct_size %u != %zu (ABI mismatch)", a->struct_size,
                     sizeof(csa_linear_args_t));
  if (a->dtype != CSA_DTYPE_F16 && a->dtype != CSA_DTYPE_BF16) return set_error(CSA_E_BADARG, "csa_linear: dtype %d", a->dtype);
  if (a->m <= 0 || a->n <= 0 || a->k <= 0 || a->ldx < a->k || a->ldw < a->k || a->ldy < a->n)
    return set_error(CSA_E_BADARG, "csa_linear: bad sizes M=%lld N=%lld K=%lld", (long long)a->m, (long long)a->n,
                     (long long)a->k);
  if (!a->x || !a->w || !a->y) return set_error(CSA_E_BADARG, "csa_linear: null pointer");
  if (((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w) | reinterpret_cast<uintptr_t>(a->y) |
        reinterpret_cast<uintptr_t>(a->bias) | reinterpret_cast<uintptr_t>(a->workspace)) & 15) ||
      ((a->ldx | a->ldw | a->ldy) & 7))
    return set_error(CSA_E_BADARG, "csa_linear: pointers must be 16-byte aligned, leading dimensions multiples of 8");
  int dev = 0;
  cudaError_t ce = cudaGetDevice(&dev);
  if (ce != cudaSuccess) return set_error(static_cast<int>(ce), "cudaGetDevice: %s", cudaGetErrorString(ce));
  if (dev < 0 || dev >= 64) return set_error(CSA_E_DEVICE, "csa_linear: device index %d", dev);
  std::lock_guard<std::mutex> lock(g_lin_mu);
  if (!g_lt[dev]) LT_TRY(cublasLtCreate(&g_lt[dev]));
  LinearKey key{dev, a->dtype, a->bias ? 1 : 0, a->m, a->n, a->k, a->ldx, a->ldw, a->ldy};
  const LinearPlan* p = nullptr;
  const size_t ws_bytes = a->workspace ? static_cast<size_t>(a->workspace_bytes) : 0;
  int rc = get_plan(key, ws_bytes, g_lt[dev], &p);
  if (rc) return rc;
  cublasLtMatmulDesc_t op = p->op;
  if (a->bias) LT_TRY(cublasLtMatmulDescSetAttribute(op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &a->bias, sizeof(a->bias)));
  const float one = 1.0f, zero = 0.0f;
  LT_TRY(cublasLtMatmul(g_lt[dev], op, &one, a->w, p->a, a->x, p->b, &zero, a->y, p->c, a->y, p->c, &p->algo,
                        a->workspace, ws_bytes, static_cast<cudaStream_t>(stream)));
  return 0;
}

extern "C" int csa_run_batch(const csa_call_t* calls, int32_t n_calls, void* stream, int32_t* failed_index) {
  if (!calls || n_calls < 0) return set_error(CSA_E_BADARG, "csa_run_batch: null calls");
  for (int i = 0; i < n_calls; ++i) {
    int rc = 0;
    const void* a = calls[i].args;
    switch (calls[i].kind) {
      case CSA_CALL_LINEAR: rc = csa_linear(static_cast<const csa_linear_args_t*>(a), stream); break;
      case CSA_CALL_GEMM: rc = csa_gemm(static_cast<const csa_gemm_args_t*>(a), stream); break;
      case CSA_CALL_ATTN: rc = csa_attn_fwd(static_cast<const csa_attn_args_t*>(a), stream); break;
      case CSA_CALL_GATHER_KV: {
        const csa_gather_kv_args_t* g = static_cast<const csa_gather_kv_args_t*>(a);
        rc = !g ? set_error(CSA_E_BADARG, "csa_run_batch: null gather_kv args")
                : csa_gather_kv(g->k, g->v, g->ld_bytes, g->group_rows, g->n_groups, g->s_idx, g->s_count, g->max_rows,
                                g->k_out, g->v_out, g->out_ld_bytes, g->out_group_rows, g->row_bytes, stream);
        break;
      }
      case CSA_CALL_PEER_SCATTER: rc = csa_peer_scatter_kv(static_cast<const csa_peer_scatter_args_t*>(a), stream); break;
      case CSA_CALL_PEER_SIGNAL: {
        const csa_peer_signal_args_t* s = static_cast<const csa_peer_signal_args_t*>(a);
        rc = !s ? set_error(CSA_E_BADARG, "csa_run_batch: null peer_signal args") : csa_peer_signal_ex(s, stream);
        break;
      }
      case CSA_CALL_EPOCH_ADVANCE: {
        const csa_epoch_advance_args_t* s = static_cast<const csa_epoch_advance_args_t*>(a);
        rc = !s ? set_error(CSA_E_BADARG, "csa_run_batch: null epoch_advance args")
                : csa_epoch_advance(s->epoch_base, s->delta, stream);
        break;
      }
      case CSA_CALL_EVENT_RECORD: {
        // on a capturing stream the record becomes an event-record NODE of the graph (external event): its time can
        // be read with cudaEventElapsedTime after a replay, which a plainly captured event does not allow
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cs);
        cudaError_t e = cudaEventRecordWithFlags(static_cast<cudaEvent_t>(const_cast<void*>(a)),
                                                 static_cast<cudaStream_t>(stream),
                                                 cs == cudaStreamCaptureStatusActive ? cudaEventRecordExternal
                                                                                     : cudaEventRecordDefault);
        rc = e == cudaSuccess ? 0 : set_error(static_cast<int>(e), "cudaEventRecord: %s", cudaGetErrorString(e));
        break;
      }
      default: rc = set_error(CSA_E_BADARG, "csa_run_batch: unknown call kind %d at index %d", calls[i].kind, i);
    }
    if (rc != 0) {
      if (failed_index) *failed_index = i;
      return rc;
    }
  }
  if (failed_index) *failed_index = -1;
  return 0;
}
