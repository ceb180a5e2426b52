// Python bindings for the msrflute_b200 sm_100a kernels.
#include <torch/extension.h>

namespace flute {
void fused_client_step(torch::Tensor w, torch::Tensor g, torch::Tensor hyper, torch::Tensor stats,
                       c10::optional<torch::Tensor> mom, c10::optional<torch::Tensor> first_step, int64_t n_logical,
                       bool nesterov, double dampening, bool zero_grad, c10::optional<torch::Tensor> prox_ref,
                       c10::optional<torch::Tensor> prox_mult, c10::optional<torch::Tensor> prox_loss);
void slot_gather_bcast(torch::Tensor W, torch::Tensor wg_slot, torch::Tensor wg, torch::Tensor map);
torch::Tensor alpha_dot(torch::Tensor wp, torch::Tensor wg, torch::Tensor gp, torch::Tensor gg, double alpha);
bool attention_supported(int64_t S, int64_t D);
std::vector<torch::Tensor> attention_fwd(torch::Tensor q, torch::Tensor k, torch::Tensor v, c10::optional<torch::Tensor> key_bias,
                                         double scale, double p_drop, torch::Tensor seed);
std::vector<torch::Tensor> attention_bwd(torch::Tensor q, torch::Tensor k, torch::Tensor v, torch::Tensor o, torch::Tensor lse,
                                         torch::Tensor d_o, c10::optional<torch::Tensor> key_bias, double scale, double p_drop,
                                         torch::Tensor seed);
torch::Tensor attention_dropout_mask(int64_t B, int64_t H, int64_t S, double p_drop, torch::Tensor seed);
void slot_pg_sqnorm(torch::Tensor W, torch::Tensor wg_slot, torch::Tensor out);
void slot_gather_fused(torch::Tensor acc_slot, torch::Tensor W, torch::Tensor wg_slot, torch::Tensor coef,
                       c10::optional<torch::Tensor> sig, c10::optional<torch::Tensor> seed);
void slot_scatter_acc(torch::Tensor acc, torch::Tensor acc_slot, torch::Tensor map);
void dead_coord_noise(torch::Tensor acc, torch::Tensor idx, torch::Tensor sig2_sum, int64_t seed);
void clip_and_stats(torch::Tensor g, torch::Tensor hyper, torch::Tensor stats, int64_t n_logical);
void fused_client_adamw(torch::Tensor w, torch::Tensor g, torch::Tensor m, torch::Tensor v, torch::Tensor step,
                        torch::Tensor hyper, torch::Tensor stats, int64_t n_logical, double beta1, double beta2, double eps,
                        bool correct_bias, bool zero_grad);
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
void sharded_server_update(std::vector<torch::Tensor> ws, std::vector<torch::Tensor> accs, std::vector<torch::Tensor> wsums,
                           c10::optional<torch::Tensor> m, c10::optional<torch::Tensor> v,
                           std::vector<torch::Tensor> m_mirror, std::vector<torch::Tensor> v_mirror, int64_t acc_mc,
                           int64_t w_mc, int64_t rank, int64_t kind, int64_t step, double lr, double b1, double b2,
                           double eps, double wd, double mom, double damp, bool nesterov, bool correct_bias,
                           double noise_scale, int64_t seed);
std::vector<torch::Tensor> group_norm_fwd(torch::Tensor x, torch::Tensor weight, torch::Tensor bias,
                                          c10::optional<torch::Tensor> residual, int64_t G, double eps, bool relu,
                                          bool per_group_affine, int64_t sets);
std::vector<torch::Tensor> group_norm_bwd(torch::Tensor dy, torch::Tensor x, torch::Tensor weight, torch::Tensor mean,
                                          torch::Tensor rstd, c10::optional<torch::Tensor> y, int64_t G, bool relu,
                                          bool per_group_affine, bool has_residual, int64_t sets);

torch::Tensor gemm_bf16_mn(torch::Tensor a, torch::Tensor b, bool a_mn, bool b_mn, bool out_fp32);
torch::Tensor gemm_bf16_tn(torch::Tensor a, torch::Tensor b, c10::optional<torch::Tensor> bias, bool relu,
                           bool out_fp32);


at::Tensor gru_cell_fwd(at::Tensor gi, at::Tensor gh, at::Tensor h);
std::vector<at::Tensor> gru_cell_bwd(at::Tensor dh, at::Tensor gi, at::Tensor gh, at::Tensor h);
std::vector<at::Tensor> lstm_cell_fwd(at::Tensor gates, at::Tensor c);
std::vector<at::Tensor> lstm_cell_bwd(at::Tensor dh, at::Tensor dc, at::Tensor gates, at::Tensor c, at::Tensor c2);

std::vector<torch::Tensor> group_norm_fwd_arena(torch::Tensor x, torch::Tensor w_arena, int64_t w_off, int64_t b_off,
                                                c10::optional<torch::Tensor> residual, int64_t G, double eps, bool relu,
                                                bool per_group_affine, int64_t sets);
std::vector<torch::Tensor> group_norm_bwd_arena(torch::Tensor dy, torch::Tensor x, torch::Tensor w_arena, int64_t w_off,
                                                torch::Tensor mean, torch::Tensor rstd, c10::optional<torch::Tensor> y,
                                                int64_t G, bool relu, bool per_group_affine, bool has_residual,
                                                int64_t sets, torch::Tensor g_arena, int64_t gw_off, int64_t gb_off);
at::Tensor slot_conv_fprop(at::Tensor x, at::Tensor w_arena, int64_t w_offset, int64_t Cout, int64_t KH, int64_t KW,
                           int64_t stride, int64_t pad, bool compact);
at::Tensor slot_conv_dgrad(at::Tensor dy, at::Tensor w_arena, int64_t w_offset, int64_t Cin, int64_t Hi, int64_t Wi,
                           int64_t KH, int64_t KW, int64_t stride, int64_t pad, bool compact);
void slot_conv_wgrad(at::Tensor x, at::Tensor dy, at::Tensor g_arena, int64_t g_offset, int64_t KH, int64_t KW,
                     int64_t stride, int64_t pad, bool compact);
void slot_conv_set_impl(int64_t impl);
void gemm_set_impl(int64_t impl);
std::vector<torch::Tensor> layer_norm_fwd(torch::Tensor x, c10::optional<torch::Tensor> weight, c10::optional<torch::Tensor> bias,
                                          double eps);
std::vector<torch::Tensor> layer_norm_bwd(torch::Tensor dy, torch::Tensor x, c10::optional<torch::Tensor> weight,
                                          torch::Tensor mean, torch::Tensor rstd);
void slot_scatter_in(torch::Tensor W, torch::Tensor wg, torch::Tensor map);
void accumulate_pseudo_grad_mapped(torch::Tensor acc, torch::Tensor wg, torch::Tensor wl, torch::Tensor weights,
                                   c10::optional<torch::Tensor> active, torch::Tensor map);
at::Tensor seg_minmax(at::Tensor flat, at::Tensor seg);
std::vector<at::Tensor> quantize_segments(at::Tensor flat, at::Tensor seg, at::Tensor stats, int64_t bits, bool emit_codes);
at::Tensor local_dp(at::Tensor flat, double max_grad, double sigma, bool clip_only, int64_t seed);
std::vector<at::Tensor> softmax_ce(at::Tensor logits, at::Tensor target, double grad_scale, int64_t ignore_index, bool want_grad);
at::Tensor cosine_stats(at::Tensor a, at::Tensor b);
std::vector<at::Tensor> max_pool2d_fwd(at::Tensor x, int64_t k, int64_t stride, int64_t pad);
at::Tensor max_pool2d_bwd(at::Tensor dy, at::Tensor arg, int64_t H, int64_t W, int64_t k, int64_t stride, int64_t pad);
void bind_slotnet(pybind11::module_& m);
torch::Tensor slot_quant_stats(torch::Tensor W, torch::Tensor wg, torch::Tensor segs, double q, int64_t bits);
void slot_quant_gather(torch::Tensor acc_slot, torch::Tensor W, torch::Tensor wg, torch::Tensor coef, torch::Tensor params,
                       torch::Tensor seg_of_blk, int64_t bits);
torch::Tensor embedding_fwd(torch::Tensor idx, torch::Tensor weight);
torch::Tensor embedding_bwd(torch::Tensor idx, torch::Tensor dy, int64_t V, int64_t padding_idx);
torch::Tensor dropout_apply(torch::Tensor x, double p, torch::Tensor seed, bool backward);
std::vector<torch::Tensor> batch_norm_fwd(torch::Tensor x, c10::optional<torch::Tensor> gamma, c10::optional<torch::Tensor> beta,
                                          c10::optional<torch::Tensor> residual, c10::optional<torch::Tensor> run_mean,
                                          c10::optional<torch::Tensor> run_var, double momentum, double eps, bool relu);
std::vector<torch::Tensor> batch_norm_bwd(torch::Tensor dy, torch::Tensor x, torch::Tensor y, c10::optional<torch::Tensor> gamma,
                                          torch::Tensor stats, bool relu, bool want_dres);
bool lstm_supported(int64_t H);
std::vector<torch::Tensor> lstm_layer_fwd(torch::Tensor gx, torch::Tensor whh, c10::optional<torch::Tensor> h0,
                                          c10::optional<torch::Tensor> c0);
std::vector<torch::Tensor> lstm_layer_bwd(torch::Tensor dhs, torch::Tensor gates, torch::Tensor cs, torch::Tensor whh,
                                          c10::optional<torch::Tensor> c0, c10::optional<torch::Tensor> dhT,
                                          c10::optional<torch::Tensor> dcT);
}  // namespace flute

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "msrflute_b200 hand-written sm_100a kernels";
  m.def("fused_client_step", &flute::fused_client_step, pybind11::arg("w"), pybind11::arg("g"), pybind11::arg("hyper"),
        pybind11::arg("stats"), pybind11::arg("mom"), pybind11::arg("first_step"), pybind11::arg("n_logical"),
        pybind11::arg("nesterov"), pybind11::arg("dampening"), pybind11::arg("zero_grad"),
        pybind11::arg("prox_ref") = pybind11::none(), pybind11::arg("prox_mult") = pybind11::none(),
        pybind11::arg("prox_loss") = pybind11::none());
  m.def("slot_gather_bcast", &flute::slot_gather_bcast);
  m.def("alpha_dot", &flute::alpha_dot);
  m.def("attention_supported", &flute::attention_supported);
  m.def("attention_fwd", &flute::attention_fwd);
  m.def("attention_bwd", &flute::attention_bwd);
  m.def("attention_dropout_mask", &flute::attention_dropout_mask);
  m.def("slot_pg_sqnorm", &flute::slot_pg_sqnorm);
  m.def("slot_gather_fused", &flute::slot_gather_fused);
  m.def("slot_scatter_acc", &flute::slot_scatter_acc);
  m.def("dead_coord_noise", &flute::dead_coord_noise);
  m.def("fused_client_adamw", &flute::fused_client_adamw);
  m.def("clip_and_stats", &flute::clip_and_stats);
  m.def("pseudo_grad", &flute::pseudo_grad);
  m.def("accumulate_pseudo_grad", &flute::accumulate_pseudo_grad);
  m.def("server_update", &flute::server_update);
  m.def("p2p_broadcast", &flute::p2p_broadcast);
  m.def("sharded_server_update", &flute::sharded_server_update);
  m.def("group_norm_fwd", &flute::group_norm_fwd);
  m.def("group_norm_bwd", &flute::group_norm_bwd);

  m.def("gemm_bf16_tn", &flute::gemm_bf16_tn);
  m.def("gemm_bf16_mn", &flute::gemm_bf16_mn);


  m.def("group_norm_fwd_arena", &flute::group_norm_fwd_arena);
  m.def("group_norm_bwd_arena", &flute::group_norm_bwd_arena);
  m.def("slot_conv_fprop", &flute::slot_conv_fprop);
  m.def("slot_conv_dgrad", &flute::slot_conv_dgrad);
  m.def("slot_conv_wgrad", &flute::slot_conv_wgrad);
  m.def("slot_conv_set_impl", &flute::slot_conv_set_impl);
  m.def("gemm_set_impl", &flute::gemm_set_impl);
  m.def("layer_norm_fwd", &flute::layer_norm_fwd);
  m.def("layer_norm_bwd", &flute::layer_norm_bwd);
  m.def("slot_scatter_in", &flute::slot_scatter_in);
  m.def("accumulate_pseudo_grad_mapped", &flute::accumulate_pseudo_grad_mapped);
  m.def("seg_minmax", &flute::seg_minmax);
  m.def("quantize_segments", &flute::quantize_segments);
  m.def("local_dp", &flute::local_dp);
  m.def("softmax_ce", &flute::softmax_ce);
  m.def("cosine_stats", &flute::cosine_stats);
  m.def("max_pool2d_fwd", &flute::max_pool2d_fwd);
  m.def("max_pool2d_bwd", &flute::max_pool2d_bwd);
  m.def("gru_cell_fwd", &flute::gru_cell_fwd);
  m.def("gru_cell_bwd", &flute::gru_cell_bwd);
  m.def("lstm_cell_fwd", &flute::lstm_cell_fwd);
  m.def("lstm_cell_bwd", &flute::lstm_cell_bwd);
  flute::bind_slotnet(m);
  m.def("slot_quant_stats", &flute::slot_quant_stats);
  m.def("slot_quant_gather", &flute::slot_quant_gather);
  m.def("embedding_fwd", &flute::embedding_fwd);
  m.def("embedding_bwd", &flute::embedding_bwd);
  m.def("dropout_apply", &flute::dropout_apply);
  m.def("batch_norm_fwd", &flute::batch_norm_fwd);
  m.def("batch_norm_bwd", &flute::batch_norm_bwd);
  m.def("lstm_supported", &flute::lstm_supported);
  m.def("lstm_layer_fwd", &flute::lstm_layer_fwd);
  m.def("lstm_layer_bwd", &flute::lstm_layer_bwd);

}
