// TEST-ONLY: compiles the kernel source (ur3e_b200/csrc/engine.cuh) as plain C++ so that its
// arithmetic can be checked against the oracle without a GPU.  Not part of the package, never
// loaded by ur3e_b200: the product has no CPU path.
#include <cstring>
#include <map>
#include <memory>
#include <string>

#include "../../ur3e_b200/csrc/compile_model.h"
#include "../../ur3e_b200/csrc/engine.cuh"
#include "../../ur3e_b200/csrc/merge_bodies.h"

using namespace ur3e;

static std::map<std::string, HostModel> g_models;
static const HostModel& get_model(const char* path) {
  auto it = g_models.find(path);
  if (it == g_models.end()) it = g_models.emplace(path, load_mjcf(path)).first;
  return it->second;
}

template <typename Real, typename D>
static int run(const HostModel& h, const double* qpos, const double* qvel, const double* ctrl, const double* ws, int nsteps, int max_iter, double tol,
               double* oq, double* ov, double* oa, double* oM, double* obias, double* ofc, int* info) {
  std::vector<int> bmap;
  DevModel<Real> m = compile_model<Real>(merge_fixed_bodies(h, bmap));
  auto s = std::make_unique<Arena<Real, D>>();
  std::memset(s.get(), 0, sizeof(Arena<Real, D>));
  for (int i = 0; i < h.nq; ++i) s->st.qpos[i] = (Real)qpos[i];
  for (int i = 0; i < h.nv; ++i) { s->st.qvel[i] = (Real)qvel[i]; s->st.qacc_ws[i] = (Real)ws[i]; }
  for (int i = 0; i < h.nu; ++i) s->ctrl[i] = (Real)ctrl[i];
  s->cap_con = D::MAXCON; s->cap_efc = D::MAXEFC;
  SolverOpts<Real> opt{max_iter, 50, (Real)tol, (Real)(sizeof(Real) == 8 ? 1e-14 : 1e-5), (Real)(sizeof(Real) == 8 ? 1e-15 : 2e-6), (Real)(sizeof(Real) == 8 ? 0.0 : 1e-8)};
  int w = 0;
  if (nsteps == 0) forward(m, *s, opt, true);
  for (int k = 0; k < nsteps; ++k) w |= substep(m, *s, opt, opt);
  for (int i = 0; i < h.nq; ++i) oq[i] = s->st.qpos[i];
  for (int i = 0; i < h.nv; ++i) { ov[i] = s->st.qvel[i]; oa[i] = nsteps == 0 ? s->qacc[i] : s->st.qacc_ws[i]; obias[i] = s->qfrc_bias[i]; ofc[i] = s->qfrc_constraint[i]; }
  for (int i = 0; i < h.nv; ++i) for (int j = 0; j < h.nv; ++j) { int hi = i > j ? i : j, lo = i > j ? j : i; oM[i * h.nv + j] = s->M[hi * (hi + 1) / 2 + lo]; }
  info[0] = s->ncon; info[1] = s->nefc; info[2] = s->solver_iter; info[3] = w; info[4] = s->overflow; info[5] = (int)sizeof(Arena<Real, D>);
  return 0;
}

extern "C" int hc_run(const char* xml, int use_float, const double* qpos, const double* qvel, const double* ctrl, const double* ws, int nsteps, int max_iter,
                      double tol, double* oq, double* ov, double* oa, double* oM, double* obias, double* ofc, int* info) {
  try {
    const HostModel& h = get_model(xml);
#define GO(Real, D) return run<Real, D>(h, qpos, qvel, ctrl, ws, nsteps, max_iter, tol, oq, ov, oa, oM, obias, ofc, info)
    if (h.nv <= DimsRaw::NV && h.npair == 0) { if (use_float) GO(float, DimsRaw); else GO(double, DimsRaw); }
    else if (h.nv <= DimsGrip::NV) { if (use_float) GO(float, DimsGrip); else GO(double, DimsGrip); }
    else { if (use_float) GO(float, DimsMain); else GO(double, DimsMain); }
  } catch (const std::exception& e) { std::fprintf(stderr, "hc_run: %s\n", e.what()); return -1; }
}

// forward pass at (qpos, qvel, ctrl) followed by the logging-sensor cold path: the [NSENSOR] record the step kernel writes
template <typename Real, typename D>
static int sensors(const HostModel& h, const double* qpos, const double* qvel, const double* ctrl, double* out) {
  std::vector<int> bmap;
  DevModel<Real> m = compile_model<Real>(merge_fixed_bodies(h, bmap));
  auto s = std::make_unique<Arena<Real, D>>();
  std::memset(s.get(), 0, sizeof(Arena<Real, D>));
  for (int i = 0; i < h.nq; ++i) s->st.qpos[i] = (Real)qpos[i];
  for (int i = 0; i < h.nv; ++i) s->st.qvel[i] = (Real)qvel[i];
  for (int i = 0; i < h.nu; ++i) s->ctrl[i] = (Real)ctrl[i];
  s->cap_con = D::MAXCON; s->cap_efc = D::MAXEFC;
  SolverOpts<Real> opt{50, 50, (Real)1e-15, (Real)1e-14, (Real)1e-15, (Real)0};
  forward(m, *s, opt, true);
  Real rec[NSENSOR];
  sensors_cold(m, *s, rec);
  for (int i = 0; i < NSENSOR; ++i) out[i] = (double)rec[i];
  return 0;
}
extern "C" int hc_sensors(const char* xml, const double* qpos, const double* qvel, const double* ctrl, double* out) {
  try {
    const HostModel& h = get_model(xml);
    if (h.nv <= DimsRaw::NV && h.npair == 0) return sensors<double, DimsRaw>(h, qpos, qvel, ctrl, out);
    if (h.nv <= DimsGrip::NV) return sensors<double, DimsGrip>(h, qpos, qvel, ctrl, out);
    return sensors<double, DimsMain>(h, qpos, qvel, ctrl, out);
  } catch (const std::exception& e) { std::fprintf(stderr, "hc_sensors: %s\n", e.what()); return -1; }
}

extern "C" int hc_model_dims(const char* xml, int* out) {
  try {
    const HostModel& h = get_model(xml);
    int v[] = {h.nq, h.nv, h.nu, h.nbody, h.njnt, h.ngeom, h.nsite, h.neq, h.ntendon, h.npair, h.nkey};
    std::memcpy(out, v, sizeof v); return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "hc_model_dims: %s\n", e.what()); return -1; }
}
extern "C" int hc_model_array(const char* xml, const char* name, double* out, int cap) {
  try {
    const HostModel& h = get_model(xml);
    auto it = h.arr.find(name); if (it == h.arr.end()) return -1;
    int n = it->second.is_int ? (int)it->second.i.size() : (int)it->second.d.size();
    for (int k = 0; k < n && k < cap; ++k) out[k] = it->second.is_int ? it->second.i[k] : it->second.d[k];
    return n;
  } catch (const std::exception& e) { return -2; }
}
