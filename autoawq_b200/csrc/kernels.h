// Internal launcher interface between the C-ABI layer (cabi.cu) and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct b200awq_op;

namespace b200awq {

struct GemmArgs {
  const void* x;        // [M, ldx] fp16
  int64_t ldx;
  const int32_t* qweight;
  const void* scales;
  const int32_t* qzeros;
  const void* bias;     // [N] fp16 or nullptr
  void* y;              // [M, N] fp16
  int M, K, N, G;
};

struct FastArgs {
  const void* x;
  int64_t ldx;
  const int16_t* qweight;  // [N/4, K]
  const void* scales;      // [8 zw, N]
  const void* szeros;      // [8 zw, N] = -z*s
  const void* bias;
  void* y;
  int M, K, N, G;
};

// workspace carve-up (see b200awq_workspace_bytes): tickets first, fp32 accumulators after
constexpr size_t kTicketBytes = 16384;  // 4096 int tickets
constexpr int kMaxSplitM = 128;         // rows of 8-byte scratch kept for split-K (= 256 rows of fp32 partial sums)

// knobs (cabi.cu)
int knob(int key);

// Kernel launch with the optional PDL attribute (knob 4).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (knob(4) != 0) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

cudaError_t dequantize_gemm(const int32_t* qweight, const void* scales, const int32_t* qzeros, void* out, int K,
                            int N, int G, cudaStream_t st);

// 2-D tiled tensor map over a row-major matrix (cached by address + shape; weights are static, activations
// recycle a few buffers).  elem_kind: 0 = fp16, 1 = int32.  128B swizzle, zero fill out of bounds.
cudaError_t make_tmap_2d(const void* ptr, int elem_kind, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                         uint32_t box_inner, uint32_t box_outer, CUtensorMap* out, bool swizzle128 = true);

bool gemv_gemm_layout_supported(const GemmArgs& a);
bool gemv_v3_supported(const GemmArgs& a);
cudaError_t gemv_v3(const GemmArgs& a, float* acc_ws, int* tickets, cudaStream_t st);
cudaError_t gemv_v3_debug_read(void* dst, size_t bytes);
cudaError_t gemv_gemm_layout(const GemmArgs& a, float* acc_ws, int* tickets, cudaStream_t st);
cudaError_t gemv_gemv_layout(const GemmArgs& a, cudaStream_t st);
cudaError_t gemv_fast_layout(const FastArgs& a, cudaStream_t st);

// tensor-core path; layout: 0 = GEMM, 1 = GEMV, 2 = FAST (qweight/qzeros reinterpretations documented in gemm_tc.cu)
cudaError_t gemm_tc(const GemmArgs& a, int layout, float* acc_ws, int* tickets, cudaStream_t st);
// true when gemm_tc(layout 0) would run the small-M kernel (TMA-staged packed weights) for these arguments
bool gemm_tcq_applicable(const GemmArgs& a, const float* acc_ws, const int* tickets);
bool gemm_tcq_shape_ok(int M, int K, int N, int G);                       // the kernel's shape envelope
int gemm_tcq_grid(int n_tiles, int KP, int M, int sms, int mode);         // its work cut (host logic, CPU-testable)
cudaError_t gemm_tcq_debug_read(void* dst, size_t bytes);   // phase timestamps of the last small-M launch (knob 3 == 9)

// decode program (program.cu)
struct Program;
int program_create(const struct ::b200awq_op* ops, int n, Program** out, cudaError_t* cuda_err);
int program_max_n(const Program* p);
int program_m(const Program* p);
int program_num_ops(const Program* p);
cudaError_t program_run(Program* p, float* acc_ws, cudaStream_t st);
int program_is_stream(const Program* p);
size_t program_stream_bytes(const Program* p);
// stream format (program_stream.cuh; oracle/stream_format.py): one-time re-layout of a GEMM-layout linear
size_t stream_format_bytes(int K, int N, int G);
bool stream_format_supported(int K, int N, int G, int mode);
cudaError_t stream_pack(const int32_t* qweight, const void* scales, const int32_t* qzeros, void* out, int K, int N, int G,
                        int mode, cudaStream_t st);
void program_destroy(Program* p);
cudaError_t program_debug_read(void* dst, size_t bytes);
cudaError_t program_abort_read(void* dst, size_t bytes);
cudaError_t stream_debug_read(void* dst, size_t bytes);
cudaError_t program_set_watchdog_seconds(int seconds);

// grouped persistent GEMV (gemv.cu): the decode-size path of grouped_gemm_forward
bool gemv_v3_moe_supported(int K, int N, int G, int hbs);
cudaError_t gemv_v3_moe(const void* x, int x_per_slot, const int32_t* qweight, const void* scales, const int32_t* qzeros,
                        const float* topk_w, const int* sorted_ids, const int* expert_ids, const int* num_post_pad,
                        void* y, int n_slots, int topk, int hbs, int E, int K, int N, int G, int block_size,
                        float* acc_ws, int* tickets, cudaStream_t st);

// MoE (moe.cu)
cudaError_t topk_softmax(const float* gating, float* topk_w, int* topk_ids, int* src_rows, int M, int E, int topk,
                         cudaStream_t st);
cudaError_t moe_align_block_size(const int* topk_ids, int numel, int num_experts, int block_size, int* sorted_ids,
                                 int* expert_ids, int* num_post_pad, cudaStream_t st);
bool moe_grouped_supported(int K, int N, int G);
cudaError_t moe_grouped_gemm(const void* x, int x_per_slot, const int32_t* qweight, const void* scales,
                             const int32_t* qzeros, const float* topk_w, const int* sorted_ids, const int* expert_ids,
                             const int* num_post_pad, void* y, int n_slots, int topk, int sorted_len, int K, int N, int G,
                             int mul_weights, int block_size, cudaStream_t st);

// one-shot all-reduce over peer memory (comm.cu)
struct Comm;
int comm_create(int rank, int world, int max_elems, Comm** out, cudaError_t* err);
cudaError_t comm_ipc_handle(Comm* c, void* out64);
cudaError_t comm_open(Comm* c, const void* handles);
cudaError_t comm_all_reduce(Comm* c, void* y, int n, cudaStream_t st);
bool comm_ready(const Comm* c);
int comm_max_elems(const Comm* c);
cudaError_t comm_error_flag(Comm* c, int* out);
void comm_destroy(Comm* c);

cudaError_t rmsnorm(const void* x, const void* w, void* out, int rows, int hidden, float eps, cudaStream_t st);
cudaError_t silu_and_mul(const void* gate_up, void* out, int rows, int d, cudaStream_t st);

}  // namespace b200awq
