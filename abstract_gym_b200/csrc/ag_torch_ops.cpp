// ag_torch_ops.cpp -- torch.ops.abstract_gym_b200.*: a compiled custom-op library over the C ABI.
//
// Thin by design (BASELINE.json north_star: "a thin C-ABI PyTorch custom-op extension"): every op checks device /
// dtype / contiguity / shapes, takes the current CUDA stream of the tensors' device and calls the extern "C" entry
// point of include/abstract_gym_b200.h in libabstract_gym_b200.so.  There is no CPU dispatch key: the ops are
// registered for CUDA only, so CPU tensors raise NotImplementedError.  State tensors are mutated in place
// (declared in the schemas), nothing is allocated and nothing synchronises, so the ops can sit inside CUDA-graph
// captures next to a policy network.
//
// Scene constants travel as a float64[13] CPU tensor (ag_params order, see pack_params in ops.py) that is read in
// place; the grid as its device tensors plus six scalars.  The reference methods each op replaces are cited in the
// header next to the entry point it forwards to.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "../../include/abstract_gym_b200.h"

namespace {

using at::Tensor;
using OptTensor = std::optional<Tensor>;

const char *status_text(ag_status s) { return ag_status_string(s); }

void check_status(ag_status s, const char *what) {
    TORCH_CHECK(s == AG_OK, what, " failed: status ", (int)s, " (", status_text(s), ")");
}

void need(const Tensor &t, at::ScalarType dt, const char *name, c10::Device dev) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (no CPU dispatch)");
    TORCH_CHECK(t.device() == dev, name, " is on ", t.device(), ", expected ", dev);
    TORCH_CHECK(t.scalar_type() == dt, name, " must be ", dt, ", got ", t.scalar_type());
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

template <typename T>
T *opt_ptr(const OptTensor &t, at::ScalarType dt, const char *name, c10::Device dev) {
    if (!t.has_value() || !t->defined()) return nullptr;
    need(*t, dt, name, dev);
    return reinterpret_cast<T *>(t->data_ptr());
}

ag_params unpack_params(const Tensor &p) {
    TORCH_CHECK(p.device().is_cpu() && p.scalar_type() == at::kDouble && p.numel() == 13 && p.is_contiguous(),
                "params must be a contiguous float64[13] CPU tensor (ops.pack_params)");
    const double *v = p.data_ptr<double>();
    ag_params q;
    q.link_1 = v[0]; q.link_2 = v[1]; q.target_x = v[2]; q.target_y = v[3]; q.target_j1 = v[4]; q.target_j2 = v[5];
    q.reach_eps = v[6]; q.section_eps = v[7]; q.reward_collision = v[8]; q.reward_reach = v[9]; q.action_scale = v[10];
    q.choose_j_tar = (int32_t)v[11]; q.max_reset_tries = (int32_t)v[12];
    return q;
}

ag_grid make_grid(const Tensor &bits, const OptTensor &bits_t, const OptTensor &hier, const Tensor &min_x, const Tensor &min_y, double side,
                  double env_size, int64_t S, int64_t n_grids, int64_t max_occupied, int64_t envs_per_grid, c10::Device dev) {
    need(bits, at::kInt, "grid_bits", dev);
    need(min_x, at::kDouble, "min_x", dev);
    need(min_y, at::kDouble, "min_y", dev);
    TORCH_CHECK(min_x.numel() >= S && min_y.numel() >= S, "corner tables shorter than S");
    ag_grid g;
    g.bits = reinterpret_cast<const uint32_t *>(bits.data_ptr());
    g.bits_t = opt_ptr<const uint32_t>(bits_t, at::kInt, "grid_bits_t", dev);
    g.hier = opt_ptr<const void>(hier, at::kByte, "grid_hier", dev);
    TORCH_CHECK(g.hier == nullptr || hier->numel() == n_grids * ag_grid_hier_bytes((int32_t)S),
                "grid_hier must hold n_grids * ag_grid_hier_bytes(S) bytes (ag_grid_pack_hier)");
    g.min_x = min_x.data_ptr<double>(); g.min_y = min_y.data_ptr<double>();
    g.side = side; g.env_size = env_size;
    g.S = (int32_t)S; g.words_per_row = ag_grid_words_per_row((int32_t)S);
    g.n_grids = (int32_t)n_grids; g.max_occupied = (int32_t)max_occupied;
    g.grid_stride_words = ag_grid_stride_words((int32_t)S);
    g.envs_per_grid = envs_per_grid;
    TORCH_CHECK(bits.numel() == n_grids * g.grid_stride_words, "grid_bits must hold n_grids * ag_grid_stride_words(S) words");
    return g;
}

void *stream_of(c10::Device dev) { return c10::cuda::getCurrentCUDAStream(dev.index()).stream(); }

// Scene.collision_check (scenario/scene_0.py:60-76) -> ag_collision_check
void collision_check(const Tensor &params, const Tensor &bits, const OptTensor &bits_t, const OptTensor &hier, const Tensor &min_x, const Tensor &min_y,
                     double side, double env_size, int64_t S, int64_t n_grids, int64_t max_occupied, int64_t envs_per_grid,
                     const Tensor &j1, const Tensor &j2, Tensor hit, OptTensor first_hit, int64_t env_id0, int64_t engine) {
    const c10::Device dev = j1.device();
    c10::cuda::CUDAGuard guard(dev);
    need(j1, at::kDouble, "j1", dev); need(j2, at::kDouble, "j2", dev); need(hit, at::kByte, "hit", dev);
    TORCH_CHECK(j2.numel() == j1.numel() && hit.numel() == j1.numel(), "j1, j2, hit must have one element per env");
    const ag_params p = unpack_params(params);
    const ag_grid g = make_grid(bits, bits_t, hier, min_x, min_y, side, env_size, S, n_grids, max_occupied, envs_per_grid, dev);
    check_status(ag_collision_check(&p, &g, j1.data_ptr<double>(), j2.data_ptr<double>(), hit.data_ptr<uint8_t>(),
                                    opt_ptr<int32_t>(first_hit, at::kInt, "first_hit", dev), j1.numel(), env_id0,
                                    (int32_t)engine, stream_of(dev)), "ag_collision_check");
}

// Scene.step (scenario/scene_0.py:88-103) -> ag_step
void step(const Tensor &params, const Tensor &bits, const OptTensor &bits_t, const OptTensor &hier, const Tensor &min_x, const Tensor &min_y, double side,
          double env_size, int64_t S, int64_t n_grids, int64_t max_occupied, int64_t envs_per_grid, Tensor j1, Tensor j2,
          const Tensor &actions, Tensor reward, Tensor flags, OptTensor ee, OptTensor dist, OptTensor first_hit, OptTensor stats,
          const OptTensor &targets, int64_t env_id0, int64_t engine) {
    const c10::Device dev = j1.device();
    c10::cuda::CUDAGuard guard(dev);
    need(j1, at::kDouble, "j1", dev); need(j2, at::kDouble, "j2", dev);
    need(reward, at::kFloat, "reward", dev); need(flags, at::kByte, "flags", dev);
    const int64_t n = j1.numel();
    TORCH_CHECK(actions.is_cuda() && actions.is_contiguous() && actions.numel() == 2 * n &&
                (actions.scalar_type() == at::kFloat || actions.scalar_type() == at::kDouble),
                "actions must be a contiguous CUDA [N,2] float32 or float64 tensor");
    const ag_params p = unpack_params(params);
    const ag_grid g = make_grid(bits, bits_t, hier, min_x, min_y, side, env_size, S, n_grids, max_occupied, envs_per_grid, dev);
    check_status(ag_step(&p, &g, j1.data_ptr<double>(), j2.data_ptr<double>(), actions.data_ptr(),
                         actions.scalar_type() == at::kFloat ? 1 : 0, reward.data_ptr<float>(), flags.data_ptr<uint8_t>(),
                         opt_ptr<double>(ee, at::kDouble, "ee", dev), opt_ptr<double>(dist, at::kDouble, "dist", dev),
                         opt_ptr<int32_t>(first_hit, at::kInt, "first_hit", dev), opt_ptr<int64_t>(stats, at::kLong, "stats", dev),
                         opt_ptr<const double>(targets, at::kDouble, "targets", dev), n, env_id0, (int32_t)engine, stream_of(dev)),
                 "ag_step");
}

// Scene.reset / random_valid_pose (scenario/scene_0.py:105-113,174-181) -> ag_reset
void reset(const Tensor &params, const Tensor &bits, const OptTensor &bits_t, const OptTensor &hier, const Tensor &min_x, const Tensor &min_y, double side,
           double env_size, int64_t S, int64_t n_grids, int64_t max_occupied, int64_t envs_per_grid, Tensor j1, Tensor j2,
           Tensor reward, Tensor flags, Tensor reset_ctr, const OptTensor &mask, const OptTensor &reset_u, OptTensor stats,
           int64_t seed, bool clear_flags, int64_t env_id0, int64_t engine) {
    const c10::Device dev = j1.device();
    c10::cuda::CUDAGuard guard(dev);
    need(j1, at::kDouble, "j1", dev); need(j2, at::kDouble, "j2", dev);
    need(reward, at::kFloat, "reward", dev); need(flags, at::kByte, "flags", dev); need(reset_ctr, at::kInt, "reset_ctr", dev);
    const int64_t n = j1.numel();
    int32_t R = 0;
    if (reset_u.has_value() && reset_u->defined()) {
        TORCH_CHECK(reset_u->dim() == 3 && reset_u->size(0) == n && reset_u->size(2) == 2, "reset_u must be [N,R,2]");
        R = (int32_t)reset_u->size(1);
    }
    const ag_params p = unpack_params(params);
    const ag_grid g = make_grid(bits, bits_t, hier, min_x, min_y, side, env_size, S, n_grids, max_occupied, envs_per_grid, dev);
    check_status(ag_reset(&p, &g, j1.data_ptr<double>(), j2.data_ptr<double>(), reward.data_ptr<float>(), flags.data_ptr<uint8_t>(),
                          reinterpret_cast<uint32_t *>(reset_ctr.data_ptr()), opt_ptr<const uint8_t>(mask, at::kByte, "mask", dev),
                          opt_ptr<const double>(reset_u, at::kDouble, "reset_u", dev), R, (uint64_t)seed, clear_flags ? 1 : 0,
                          opt_ptr<int64_t>(stats, at::kLong, "stats", dev), n, env_id0, (int32_t)engine, stream_of(dev)),
                 "ag_reset");
}

// the loop body of experiment/experiment_0.py:20-34 fused over K steps -> ag_rollout
void rollout(const Tensor &params, const Tensor &bits, const OptTensor &bits_t, const OptTensor &hier, const Tensor &min_x, const Tensor &min_y, double side,
             double env_size, int64_t S, int64_t n_grids, int64_t max_occupied, int64_t envs_per_grid, Tensor j1, Tensor j2,
             Tensor reward, Tensor flags, Tensor step_ctr, Tensor reset_ctr, Tensor ep_len, const OptTensor &actions,
             const OptTensor &reset_u, const OptTensor &targets, OptTensor rec_j1, OptTensor rec_j2, OptTensor rec_reward,
             OptTensor rec_flags, Tensor stats, OptTensor diag, OptTensor events, OptTensor event_count, int64_t K, int64_t seed,
             int64_t env_id0, int64_t engine) {
    const c10::Device dev = j1.device();
    c10::cuda::CUDAGuard guard(dev);
    need(j1, at::kDouble, "j1", dev); need(j2, at::kDouble, "j2", dev);
    need(reward, at::kFloat, "reward", dev); need(flags, at::kByte, "flags", dev);
    need(step_ctr, at::kInt, "step_ctr", dev); need(reset_ctr, at::kInt, "reset_ctr", dev); need(ep_len, at::kInt, "ep_len", dev);
    need(stats, at::kLong, "stats", dev);
    const int64_t n = j1.numel();
    ag_rollout_args a = {};
    a.n = n; a.env_id0 = env_id0; a.K = (int32_t)K; a.engine = (int32_t)engine; a.seed = (uint64_t)seed;
    if (actions.has_value() && actions->defined())
        TORCH_CHECK(actions->dim() == 3 && actions->size(0) == K && actions->size(1) == n && actions->size(2) == 2, "actions must be [K,N,2]");
    a.actions = opt_ptr<const float>(actions, at::kFloat, "actions", dev);
    if (reset_u.has_value() && reset_u->defined()) {
        TORCH_CHECK(reset_u->dim() == 3 && reset_u->size(0) == n && reset_u->size(2) == 2, "reset_u must be [N,R,2]");
        a.R = (int32_t)reset_u->size(1);
    }
    a.reset_u = opt_ptr<const double>(reset_u, at::kDouble, "reset_u", dev);
    a.targets = opt_ptr<const double>(targets, at::kDouble, "targets", dev);
    a.j1 = j1.data_ptr<double>(); a.j2 = j2.data_ptr<double>(); a.reward = reward.data_ptr<float>(); a.flags = flags.data_ptr<uint8_t>();
    a.step_ctr = reinterpret_cast<uint32_t *>(step_ctr.data_ptr()); a.reset_ctr = reinterpret_cast<uint32_t *>(reset_ctr.data_ptr());
    a.ep_len = reinterpret_cast<uint32_t *>(ep_len.data_ptr());
    for (const OptTensor *r : {&rec_j1, &rec_j2, &rec_reward, &rec_flags})
        if (r->has_value() && (*r)->defined()) TORCH_CHECK((*r)->numel() == K * n, "record planes must be [K,N]");
    a.rec_j1 = opt_ptr<float>(rec_j1, at::kFloat, "rec_j1", dev); a.rec_j2 = opt_ptr<float>(rec_j2, at::kFloat, "rec_j2", dev);
    a.rec_reward = opt_ptr<float>(rec_reward, at::kFloat, "rec_reward", dev);
    a.rec_flags = opt_ptr<uint8_t>(rec_flags, at::kByte, "rec_flags", dev);
    a.stats = stats.data_ptr<int64_t>();
    a.diag = opt_ptr<int64_t>(diag, at::kLong, "diag", dev);
    a.events = opt_ptr<uint32_t>(events, at::kInt, "events", dev);
    a.event_count = opt_ptr<int64_t>(event_count, at::kLong, "event_count", dev);
    if (a.events) a.event_capacity = events->numel() / 3;
    const ag_params p = unpack_params(params);
    const ag_grid g = make_grid(bits, bits_t, hier, min_x, min_y, side, env_size, S, n_grids, max_occupied, envs_per_grid, dev);
    check_status(ag_rollout(&p, &g, &a, stream_of(dev)), "ag_rollout");
}

// fused gym-style step (scenario/vector_env.py): Scene.step + same-step auto-reset + observations -> ag_step_obs
void step_obs(const Tensor &params, const Tensor &bits, const OptTensor &bits_t, const OptTensor &hier, const Tensor &min_x, const Tensor &min_y, double side,
              double env_size, int64_t S, int64_t n_grids, int64_t max_occupied, int64_t envs_per_grid, Tensor j1, Tensor j2,
              const Tensor &actions, Tensor reward, Tensor flags, Tensor reset_ctr, Tensor ep_len, const OptTensor &targets, Tensor obs,
              Tensor reward_out, Tensor terminated, Tensor collision, OptTensor final_obs, OptTensor crop, Tensor stats, int64_t seed,
              bool auto_reset, int64_t env_id0, int64_t engine) {
    const c10::Device dev = j1.device();
    c10::cuda::CUDAGuard guard(dev);
    need(j1, at::kDouble, "j1", dev); need(j2, at::kDouble, "j2", dev);
    need(reward, at::kFloat, "reward", dev); need(flags, at::kByte, "flags", dev);
    need(reset_ctr, at::kInt, "reset_ctr", dev); need(ep_len, at::kInt, "ep_len", dev);
    need(obs, at::kDouble, "obs", dev); need(reward_out, at::kFloat, "reward_out", dev);
    need(terminated, at::kByte, "terminated", dev); need(collision, at::kByte, "collision", dev); need(stats, at::kLong, "stats", dev);
    const int64_t n = j1.numel();
    TORCH_CHECK(actions.is_cuda() && actions.is_contiguous() && actions.numel() == 2 * n &&
                (actions.scalar_type() == at::kFloat || actions.scalar_type() == at::kDouble),
                "actions must be a contiguous CUDA [N,2] float32 or float64 tensor");
    TORCH_CHECK(obs.numel() == n * AG_OBS_DIM, "obs must be [N,", AG_OBS_DIM, "]");
    int32_t crop_size = 0;
    if (crop.has_value() && crop->defined()) {
        TORCH_CHECK(crop->dim() == 3 && crop->size(0) == n && crop->size(1) == crop->size(2), "crop must be [N,c,c]");
        crop_size = (int32_t)crop->size(1);
    }
    const ag_params p = unpack_params(params);
    const ag_grid g = make_grid(bits, bits_t, hier, min_x, min_y, side, env_size, S, n_grids, max_occupied, envs_per_grid, dev);
    check_status(ag_step_obs(&p, &g, j1.data_ptr<double>(), j2.data_ptr<double>(), actions.data_ptr(),
                             actions.scalar_type() == at::kFloat ? 1 : 0, reward.data_ptr<float>(), flags.data_ptr<uint8_t>(),
                             reinterpret_cast<uint32_t *>(reset_ctr.data_ptr()), reinterpret_cast<uint32_t *>(ep_len.data_ptr()),
                             opt_ptr<const double>(targets, at::kDouble, "targets", dev), obs.data_ptr<double>(),
                             reward_out.data_ptr<float>(), terminated.data_ptr<uint8_t>(), collision.data_ptr<uint8_t>(),
                             opt_ptr<double>(final_obs, at::kDouble, "final_obs", dev), opt_ptr<uint8_t>(crop, at::kByte, "crop", dev),
                             crop_size, stats.data_ptr<int64_t>(), (uint64_t)seed, auto_reset ? 1 : 0, n, env_id0, (int32_t)engine,
                             stream_of(dev)),
                 "ag_step_obs");
}

#define AG_GRID_SCHEMA "Tensor params, Tensor grid_bits, Tensor? grid_bits_t, Tensor? grid_hier, Tensor min_x, Tensor min_y, float side, float env_size, " \
                       "int S, int n_grids, int max_occupied, int envs_per_grid, "

}  // namespace

TORCH_LIBRARY(abstract_gym_b200, m) {
    m.def("collision_check(" AG_GRID_SCHEMA "Tensor j1, Tensor j2, Tensor(a!) hit, Tensor(b!)? first_hit, int env_id0, int engine) -> ()");
    m.def("step(" AG_GRID_SCHEMA "Tensor(a!) j1, Tensor(b!) j2, Tensor actions, Tensor(c!) reward, Tensor(d!) flags, Tensor(e!)? ee, "
          "Tensor(f!)? dist, Tensor(g!)? first_hit, Tensor(h!)? stats, Tensor? targets, int env_id0, int engine) -> ()");
    m.def("reset(" AG_GRID_SCHEMA "Tensor(a!) j1, Tensor(b!) j2, Tensor(c!) reward, Tensor(d!) flags, Tensor(e!) reset_ctr, Tensor? mask, "
          "Tensor? reset_u, Tensor(f!)? stats, int seed, bool clear_flags, int env_id0, int engine) -> ()");
    m.def("rollout(" AG_GRID_SCHEMA "Tensor(a!) j1, Tensor(b!) j2, Tensor(c!) reward, Tensor(d!) flags, Tensor(e!) step_ctr, "
          "Tensor(f!) reset_ctr, Tensor(g!) ep_len, Tensor? actions, Tensor? reset_u, Tensor? targets, Tensor(h!)? rec_j1, "
          "Tensor(i!)? rec_j2, Tensor(j!)? rec_reward, Tensor(k!)? rec_flags, Tensor(l!) stats, Tensor(m!)? diag, Tensor(n!)? events, "
          "Tensor(o!)? event_count, int K, int seed, int env_id0, int engine) -> ()");
    m.def("step_obs(" AG_GRID_SCHEMA "Tensor(a!) j1, Tensor(b!) j2, Tensor actions, Tensor(c!) reward, Tensor(d!) flags, "
          "Tensor(e!) reset_ctr, Tensor(f!) ep_len, Tensor? targets, Tensor(g!) obs, Tensor(h!) reward_out, Tensor(i!) terminated, "
          "Tensor(j!) collision, Tensor(k!)? final_obs, Tensor(l!)? crop, Tensor(m!) stats, int seed, bool auto_reset, int env_id0, "
          "int engine) -> ()");
}

TORCH_LIBRARY_IMPL(abstract_gym_b200, CUDA, m) {
    m.impl("collision_check", &collision_check);
    m.impl("step", &step);
    m.impl("reset", &reset);
    m.impl("rollout", &rollout);
    m.impl("step_obs", &step_obs);
}
