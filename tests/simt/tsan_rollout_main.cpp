// TEST INFRASTRUCTURE: race detector run of the fused rollout kernels.  The product's main translation unit, compiled as C++ through
// tests/simt/cuda_runtime.h (warp lanes are real threads: a warp-level exchange that is not followed by the barrier the kernel needs
// is a data race) and built with -fsanitize=thread, evaluates a case written by tests/test_simt_emulation_cpu.py:
//   file = int32 n_sections, then per section: char name[24], int64 nbytes, data padded to 16 bytes.
#include <cstdio>
#include <map>
#include <string>

#include "simt_kernels.cpp"

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::map<std::string, std::vector<char>> sec;
  int32_t n = 0;
  if (fread(&n, 4, 1, f) != 1) return 2;
  for (int i = 0; i < n; ++i) {
    char name[24];
    long long nbytes = 0;
    if (fread(name, 1, 24, f) != 24 || fread(&nbytes, 8, 1, f) != 1) return 2;
    std::vector<char> d((size_t)((nbytes + 15) / 16 * 16));
    if (!d.empty() && fread(d.data(), 1, d.size(), f) != d.size()) return 2;
    sec[std::string(name)] = std::move(d);
  }
  fclose(f);
  auto ptr = [&](const char *k) -> const void * { auto it = sec.find(k); return it == sec.end() || it->second.empty() ? nullptr : it->second.data(); };
  const int32_t *dims = (const int32_t *)ptr("dims");  // B, H, D, S, L, cub_max_n, cub_envs, vox_max_n, vox_envs, n_vox
  const int B = dims[0], H = dims[1], D = dims[2], S = dims[3], L = dims[4];
  cb200_rollout_cfg cfg;
  std::memcpy(&cfg, ptr("cfg"), sizeof(cfg));
  cb200_cuboid_set cs{};
  cb200_voxel_set vs{};
  cb200_rollout_io io{};
  io.q = (const float *)ptr("q");
  io.vel = (const float *)ptr("vel"), io.acc = (const float *)ptr("acc"), io.jerk = (const float *)ptr("jerk"), io.dt = (const float *)ptr("dt");
  io.robot_blob = ptr("blob"), io.robot_blob_host = ptr("blob"), io.robot_blob_bytes = (int32_t)sec["blob"].size();
  if (ptr("cub_dims")) {
    cs.dims = (const float *)ptr("cub_dims"), cs.inv_pose = (const float *)ptr("cub_inv_pose");
    cs.enable = (const uint8_t *)ptr("cub_enable"), cs.count = (const int32_t *)ptr("cub_count");
    cs.max_n = dims[5], cs.num_envs = dims[6];
    io.cuboids = &cs;
  }
  if (ptr("vox_params")) {
    vs.params = (const float *)ptr("vox_params"), vs.inv_pose = (const float *)ptr("vox_inv_pose");
    vs.enable = (const uint8_t *)ptr("vox_enable"), vs.count = (const int32_t *)ptr("vox_count");
    vs.features = (const uint16_t *)ptr("vox_features");
    vs.n_voxels_per_layer = dims[9], vs.max_n = dims[7], vs.num_envs = dims[8];
    vs.max_dist = *(const float *)ptr("vox_max_dist");
    io.voxels = &vs;
  }
  io.goal_position = (const float *)ptr("goal_pos"), io.goal_quat = (const float *)ptr("goal_quat");
  io.idxs_goal = (const int32_t *)ptr("idxs_goal");
  io.pose_axes_non_terminal = (const float *)ptr("axes_nt");
  const size_t N = (size_t)B * H;
  std::vector<float> cost(N), grad(N * D), selfc(N), scene(N * S), pose(N * 2 * L), csp(N * D), gv(N * D), ga(N * D), gj(N * D);
  io.cost = cost.data(), io.grad_q = grad.data(), io.self_cost = selfc.data(), io.scene_cost = scene.data();
  io.pose_cost = pose.data(), io.cspace_cost = csp.data();
  if (io.vel) io.grad_vel = gv.data(), io.grad_acc = ga.data(), io.grad_jerk = gj.data();
  io.batch_size = B, io.horizon = H;
  const int err = cb200_rollout_cost_grad(&cfg, &io, nullptr);
  if (err) { printf("error %d\n", err); return 3; }
  double sc = 0, sg = 0;
  for (float v : cost) sc += v;
  for (float v : grad) sg += std::fabs(v);
  printf("ok %.9g %.9g\n", sc, sg);
  return 0;
}
