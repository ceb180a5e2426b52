// Python bindings for the msrflute_b200 sm_100a kernels.
#include <torch/extension.h>

namespace flute {
void fused_client_step(torch::Tensor w, torch::Tensor g, torch::Tensor hyper, torch::Tensor stats,
                       c10::optional<torch::Tensor> mom, c10::optional<torch::Tensor> first_step, int64_t n_logical,
                       bool nesterov, double dampening, bool zero_grad);
void clip_and_stats(torch::Tensor g, torch::Tensor hyper, torch::Tensor stats, int64_t n_logical);
void pseudo_grad(torch::Tensor wg, torch::Tensor wl, torch::Tensor out, c10::optional<torch::Tensor> weight,
                 c10::optional<torch::Tensor> stats);
void accumulate_pseudo_grad(torch::Tensor acc, torch::Tensor wg, torch::Tensor wl, torch::Tensor weights,
                            c10::optional<torch::Tensor> active);
void server_update(torch::Tensor w, std::vector<torch::Tensor> accs, torch::Tensor weight_sum,
                   c10::optional<torch::Tensor> m, c10::optional<torch::Tensor> v, c10::optional<torch::Tensor> grad_out,
                   c10::optional<torch::Tensor> segments, std::vector<torch::Tensor> bcast,
                   c10::optional<torch::Tensor> stats_out, int64_t kind, int64_t step, double lr, double b1, double b2,
                   double eps, double wd, double mom, double damp, bool nesterov, bool correct_bias, double noise_scale,
                   int64_t seed, double max_grad_norm, bool zero_accs);
void p2p_broadcast(torch::Tensor src, std::vector<torch::Tensor> dsts);
std::vector<torch::Tensor> group_norm_fwd(torch::Tensor x, torch::Tensor weight, torch::Tensor bias,
                                          c10::optional<torch::Tensor> residual, int64_t G, double eps, bool relu,
                                          bool per_group_affine);
std::vector<torch::Tensor> group_norm_bwd(torch::Tensor dy, torch::Tensor x, torch::Tensor weight, torch::Tensor mean,
                                          torch::Tensor rstd, c10::optional<torch::Tensor> y, int64_t G, bool relu,
                                          bool per_group_affine, bool has_residual);
#ifdef FLUTE_WITH_GEMM
torch::Tensor gemm_bf16_tn(torch::Tensor a, torch::Tensor b, c10::optional<torch::Tensor> bias, bool relu,
                           bool out_fp32);
#endif
}  // namespace flute

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "msrflute_b200 hand-written sm_100a kernels";
  m.def("fused_client_step", &flute::fused_client_step);
  m.def("clip_and_stats", &flute::clip_and_stats);
  m.def("pseudo_grad", &flute::pseudo_grad);
  m.def("accumulate_pseudo_grad", &flute::accumulate_pseudo_grad);
  m.def("server_update", &flute::server_update);
  m.def("p2p_broadcast", &flute::p2p_broadcast);
  m.def("group_norm_fwd", &flute::group_norm_fwd);
  m.def("group_norm_bwd", &flute::group_norm_bwd);
#ifdef FLUTE_WITH_GEMM
  m.def("gemm_bf16_tn", &flute::gemm_bf16_tn);
#endif
}
