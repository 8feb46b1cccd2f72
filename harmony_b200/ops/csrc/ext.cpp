// pybind11 bindings of the harmony_b200 native runtime (sm_100a kernels + peer arena).
#include <torch/extension.h>
#include <string>
#include <vector>

namespace hb {
// arena.cpp
int64_t arena_create(int64_t device, int64_t bytes, int64_t kind, int64_t align);
// vmm.cpp
std::vector<bool> vmm_supported(int64_t device);
int64_t vmm_granularity(int64_t device);
int64_t vmm_export_fd(int64_t base);
int64_t vmm_import_fd(int64_t fd, int64_t bytes, int64_t device);
void vmm_grant(int64_t base, int64_t device);
void vmm_release(int64_t base);
int64_t mc_granularity(int64_t world, int64_t bytes, bool recommended);
std::vector<int64_t> mc_create(int64_t device, int64_t world, int64_t bytes);
int64_t mc_import(int64_t fd, int64_t device, int64_t world, int64_t bytes);
void mc_add_device(int64_t id, int64_t device);
void mc_bind(int64_t id, int64_t mc_offset, int64_t seg_base, int64_t mem_offset, int64_t bytes);
int64_t mc_map(int64_t id, int64_t device);
void mc_unmap(int64_t va, int64_t bytes);
void mc_destroy(int64_t id, int64_t device);
void close_fd(int64_t fd);
int64_t arena_alloc(int64_t id, int64_t bytes);
void set_sm_share(double share);
double get_sm_share();
int sm_count();
void arena_free(int64_t id, int64_t ptr);
std::vector<int64_t> arena_stats(int64_t id);
void arena_destroy(int64_t id);
py::bytes ipc_export(int64_t ptr);
int64_t ipc_open(std::string handle_bytes, int64_t device);
void ipc_close(int64_t ptr);
bool enable_peer_access(int64_t device, int64_t peer);
at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, std::vector<int64_t> strides,
                           int64_t dtype_code, int64_t device);
// table_ops.cu
void table_pull(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
                at::Tensor keys, at::Tensor out);
void table_push(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
                at::Tensor keys, at::Tensor deltas, int64_t fn, double step, double lo, double hi, double eps);
void table_push_sparse_i32(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim,
                           int64_t mode, at::Tensor keys, at::Tensor topics, at::Tensor deltas, int64_t K,
                           int64_t bitmap_off);
void table_pull_lda_sparse(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim,
                           int64_t mode, at::Tensor keys, at::Tensor out, int64_t K, int64_t bitmap_off,
                           c10::optional<at::Tensor> chunk_count);
void lda_rebuild_bitmaps(at::Tensor slab_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t K, int64_t bitmap_off);
void table_push_scalar_i32(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim,
                           int64_t mode, at::Tensor keys, at::Tensor values, int64_t op);
void init_slab(int64_t slab_ptr, int64_t rows, int64_t row_stride, int64_t row_dim, int64_t kind, double a,
               double b, int64_t seed, int64_t row_key_base, int64_t key_stride);
void migrate_copy(int64_t src_ptr, int64_t dst_ptr, int64_t nbytes);
void flip_owner(at::Tensor ptr_tables, int64_t block, at::Tensor new_ptrs);
void migrate_bulk(at::Tensor src_ptrs, at::Tensor dst_ptrs, c10::optional<at::Tensor> snap_ptrs, int64_t block_bytes,
                  at::Tensor err);
void migrate_drain(at::Tensor old_ptrs, at::Tensor snap_ptrs, at::Tensor new_ptrs, int64_t block_bytes, bool is_int,
                   at::Tensor forwarded, bool do_clamp, double lo, double hi);
void flip_blocks(at::Tensor ptr_table, at::Tensor blocks, at::Tensor new_ptrs);
void scale_blocks(at::Tensor ptrs, int64_t block_bytes, double factor);
void table_put_i32(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
                   at::Tensor keys, at::Tensor values, int64_t bitmap_off);
void clock_publish(at::Tensor clock_ptrs, int64_t rank, int64_t value);
void clock_wait(at::Tensor my_clocks, int64_t min_value, int64_t timeout_cycles, at::Tensor err_flag);
// mlr_ops.cu
void mlr_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                 int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, at::Tensor labels,
                 double step, c10::optional<at::Tensor> D, at::Tensor stats, double logit_scale);
void mlr_backward_push(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                       int64_t mode, at::Tensor uniq_feat, at::Tensor feat_ptr, at::Tensor samp,
                       at::Tensor xval, at::Tensor D, bool local_rmw, int64_t wait_ptr, int64_t wait_target,
                       c10::optional<at::Tensor> err);
// mlr_nvls_ops.cu
void ship_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                  int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, int64_t world, int64_t B,
                  int64_t tile_rows, int64_t p_ptr, int64_t tile_done_ptr, c10::optional<at::Tensor> pflag_ptrs,
                  int64_t epoch);
void ship_softmax(int64_t src_mc, c10::optional<at::Tensor> src_ptrs, int64_t src_row0, int64_t dst_mc,
                  c10::optional<at::Tensor> dst_ptrs, int64_t dst_row0, int64_t stride, at::Tensor labels,
                  int64_t num_classes, double step, double logit_scale, int64_t tile_rows, int64_t pflag_ptr,
                  int64_t pflag_target, c10::optional<at::Tensor> dcount_ptrs, at::Tensor stats, at::Tensor err,
                  bool zero_src);
void ship3_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                   int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, int64_t world, int64_t B,
                   int64_t tile_rows, at::Tensor logit_ptrs, int64_t tile_done_ptr,
                   c10::optional<at::Tensor> pflag_ptrs, int64_t epoch);
void wait_count(int64_t ptr, int64_t target, at::Tensor err);
void ship_preload(int64_t device);
// nmf_ops.cu
void nmf_batch(at::Tensor csr_ptr, at::Tensor col_slot, at::Tensor val, at::Tensor L, at::Tensor l_row,
               at::Tensor Rst, at::Tensor Gst, int64_t rank, double step, double lambda, double lo, double hi,
               at::Tensor loss);
// lda_ops.cu
void lda_sample(at::Tensor doc_ptr, at::Tensor tok_slot, at::Tensor slot_key, at::Tensor z, at::Tensor nwk,
                at::Tensor nk, int64_t K, double alpha, double beta, int64_t V, int64_t seed, int64_t tok_base,
                at::Tensor out_keys, at::Tensor out_topics, at::Tensor out_deltas, at::Tensor sum_delta,
                bool init_mode);
// mlr_ship_ops.cu
void mlr_partial_forward(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t num_classes,
                         int64_t mode, at::Tensor csr_ptr, at::Tensor csr_col, at::Tensor csr_val, int64_t B,
                         at::Tensor logit_ptrs);
void mlr_softmax_bcast(at::Tensor L, at::Tensor labels, int64_t num_classes, double step, at::Tensor d_ptrs,
                       int64_t rank, at::Tensor stats);
// mlr_dense_tc.cu
at::Tensor make_block_tensormaps(at::Tensor block_ptrs_cpu, int64_t rows_per_block, int64_t row_stride,
                                 int64_t device, int64_t swizzle_mode);
void dense_logits_tf32(at::Tensor X, at::Tensor bmaps, int64_t rows_per_block, int64_t n_cols, at::Tensor out,
                       bool use_2cta, bool stream_k);
void dense_grad_push_tf32(at::Tensor X, at::Tensor D, at::Tensor block_ptrs, int64_t rows_per_block,
                          int64_t row_stride, int64_t n_classes, int64_t mode, bool use_2cta, bool stream_k);
// k12_ops.cu
void lasso_cd(at::Tensor block_ptrs, int64_t rows_per_block, int64_t row_stride, int64_t row_dim, int64_t mode,
              at::Tensor Xt, at::Tensor y, double lambda, double step, at::Tensor w_new, at::Tensor w_old,
              at::Tensor stats);
void gbt_level(at::Tensor X, at::Tensor order, at::Tensor ftype, at::Tensor node_of, at::Tensor g,
               int64_t node_base, int64_t num_nodes, int64_t leaf_min, double lam, double gamma, double step,
               bool allow_split, at::Tensor t_feature, at::Tensor t_thr, at::Tensor t_mask, at::Tensor t_value,
               at::Tensor t_leaf);
void gbt_predict(at::Tensor X, at::Tensor t_feature, at::Tensor t_thr, at::Tensor t_mask, at::Tensor t_value,
                 at::Tensor t_leaf, at::Tensor out);
double lgamma_sum_i32(at::Tensor x, int64_t cols, double c);
// loader.cpp (host runtime)
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> parse_sparse_text(const std::string& text, int64_t mode,
                                                                              int64_t threads);
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> group_by_feature(at::Tensor csr_ptr, at::Tensor csr_col,
                                                                             at::Tensor csr_val);
}  // namespace hb

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.doc() = "harmony_b200 native runtime (sm_100a)";
  m.def("arena_create", &hb::arena_create, py::arg("device"), py::arg("bytes"), py::arg("kind") = 0,
        py::arg("align") = 0);
  m.def("vmm_supported", &hb::vmm_supported);
  m.def("vmm_granularity", &hb::vmm_granularity);
  m.def("vmm_export_fd", &hb::vmm_export_fd);
  m.def("vmm_import_fd", &hb::vmm_import_fd);
  m.def("vmm_grant", &hb::vmm_grant);
  m.def("vmm_release", &hb::vmm_release);
  m.def("mc_granularity", &hb::mc_granularity);
  m.def("mc_create", &hb::mc_create);
  m.def("mc_import", &hb::mc_import);
  m.def("mc_add_device", &hb::mc_add_device);
  m.def("mc_bind", &hb::mc_bind);
  m.def("mc_map", &hb::mc_map);
  m.def("mc_unmap", &hb::mc_unmap);
  m.def("mc_destroy", &hb::mc_destroy);
  m.def("close_fd", &hb::close_fd);
  m.def("arena_alloc", &hb::arena_alloc);
  m.def("set_sm_share", &hb::set_sm_share);
  m.def("get_sm_share", &hb::get_sm_share);
  m.def("sm_count", &hb::sm_count);
  m.def("arena_free", &hb::arena_free);
  m.def("arena_stats", &hb::arena_stats);
  m.def("arena_destroy", &hb::arena_destroy);
  m.def("ipc_export", &hb::ipc_export);
  m.def("ipc_open", &hb::ipc_open);
  m.def("ipc_close", &hb::ipc_close);
  m.def("enable_peer_access", &hb::enable_peer_access);
  m.def("tensor_from_ptr", &hb::tensor_from_ptr);
  m.def("table_pull", &hb::table_pull);
  m.def("table_push", &hb::table_push);
  m.def("table_push_sparse_i32", &hb::table_push_sparse_i32, py::arg("block_ptrs"), py::arg("rows_per_block"),
        py::arg("row_stride"), py::arg("row_dim"), py::arg("mode"), py::arg("keys"), py::arg("topics"), py::arg("deltas"),
        py::arg("K"), py::arg("bitmap_off") = 0);
  m.def("table_pull_lda_sparse", &hb::table_pull_lda_sparse, py::arg("block_ptrs"), py::arg("rows_per_block"),
        py::arg("row_stride"), py::arg("row_dim"), py::arg("mode"), py::arg("keys"), py::arg("out"), py::arg("K"),
        py::arg("bitmap_off"), py::arg("chunk_count") = py::none());
  m.def("lda_rebuild_bitmaps", &hb::lda_rebuild_bitmaps);
  m.def("table_push_scalar_i32", &hb::table_push_scalar_i32);
  m.def("init_slab", &hb::init_slab);
  m.def("migrate_copy", &hb::migrate_copy);
  m.def("flip_owner", &hb::flip_owner);
  m.def("migrate_bulk", &hb::migrate_bulk);
  m.def("migrate_drain", &hb::migrate_drain, py::arg("old_ptrs"), py::arg("snap_ptrs"), py::arg("new_ptrs"),
        py::arg("block_bytes"), py::arg("is_int"), py::arg("forwarded"), py::arg("do_clamp") = false,
        py::arg("lo") = 0.0, py::arg("hi") = 0.0);
  m.def("flip_blocks", &hb::flip_blocks);
  m.def("scale_blocks", &hb::scale_blocks);
  m.def("table_put_i32", &hb::table_put_i32, py::arg("block_ptrs"), py::arg("rows_per_block"), py::arg("row_stride"),
        py::arg("row_dim"), py::arg("mode"), py::arg("keys"), py::arg("values"), py::arg("bitmap_off") = 0);
  m.def("clock_publish", &hb::clock_publish);
  m.def("clock_wait", &hb::clock_wait);
  m.def("mlr_forward", &hb::mlr_forward, py::arg("block_ptrs"), py::arg("rows_per_block"), py::arg("row_stride"),
        py::arg("num_classes"), py::arg("mode"), py::arg("csr_ptr"), py::arg("csr_col"), py::arg("csr_val"),
        py::arg("labels"), py::arg("step"), py::arg("D"), py::arg("stats"), py::arg("logit_scale") = 1.0);
  m.def("mlr_backward_push", &hb::mlr_backward_push, py::arg("block_ptrs"), py::arg("rows_per_block"),
        py::arg("row_stride"), py::arg("num_classes"), py::arg("mode"), py::arg("uniq_feat"), py::arg("feat_ptr"),
        py::arg("samp"), py::arg("xval"), py::arg("D"), py::arg("local_rmw") = false, py::arg("wait_ptr") = 0,
        py::arg("wait_target") = 0, py::arg("err") = py::none());
  m.def("ship_forward", &hb::ship_forward);
  m.def("ship_softmax", &hb::ship_softmax, py::arg("src_mc"), py::arg("src_ptrs"), py::arg("src_row0"),
        py::arg("dst_mc"), py::arg("dst_ptrs"), py::arg("dst_row0"), py::arg("stride"), py::arg("labels"),
        py::arg("num_classes"), py::arg("step"), py::arg("logit_scale"), py::arg("tile_rows"), py::arg("pflag_ptr"),
        py::arg("pflag_target"), py::arg("dcount_ptrs"), py::arg("stats"), py::arg("err"), py::arg("zero_src") = false);
  m.def("ship3_forward", &hb::ship3_forward);
  m.def("wait_count", &hb::wait_count);
  m.def("ship_preload", &hb::ship_preload);
  m.def("nmf_batch", &hb::nmf_batch);
  m.def("lda_sample", &hb::lda_sample);
  m.def("mlr_partial_forward", &hb::mlr_partial_forward);
  m.def("dense_grad_push_tf32", &hb::dense_grad_push_tf32, py::arg("X"), py::arg("D"), py::arg("block_ptrs"),
        py::arg("rows_per_block"), py::arg("row_stride"), py::arg("n_classes"), py::arg("mode"),
        py::arg("use_2cta") = false, py::arg("stream_k") = false);
  m.def("make_block_tensormaps", &hb::make_block_tensormaps, py::arg("block_ptrs_cpu"), py::arg("rows_per_block"),
        py::arg("row_stride"), py::arg("device"), py::arg("swizzle_mode") = 1);
  m.def("dense_logits_tf32", &hb::dense_logits_tf32, py::arg("X"), py::arg("bmaps"), py::arg("rows_per_block"),
        py::arg("n_cols"), py::arg("out"), py::arg("use_2cta") = false, py::arg("stream_k") = false);
  m.def("mlr_softmax_bcast", &hb::mlr_softmax_bcast);
  m.def("lasso_cd", &hb::lasso_cd);
  m.def("gbt_level", &hb::gbt_level);
  m.def("gbt_predict", &hb::gbt_predict);
  m.def("lgamma_sum_i32", &hb::lgamma_sum_i32);
  m.def("parse_sparse_text", &hb::parse_sparse_text, py::arg("text"), py::arg("mode"), py::arg("threads") = 8);
  m.def("group_by_feature", &hb::group_by_feature);
}
