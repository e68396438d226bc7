// Model-compiler pass: rigidly attached (jointless) bodies of a moving tree are merged into their nearest jointed ancestor.
// robotiq_base_mount + gripper_base fold into wrist_3_link, the pads + silicone pads into the followers: main.xml goes from
// 25 bodies / 14 tree levels to 19 / 10, which shortens every level-serial phase of the kernel.  The merged body gets the
// composite mass, centre of mass and principal inertia of its members; geoms, sites, joints and connect anchors are
// re-expressed in the merged frame.  Constraint weights (geom_invweight0, eq_invweight0) were taken from the unmerged model.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "host_model.h"

namespace ur3e {
namespace merge_detail {
inline void qmul(double* r, const double* a, const double* b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  std::memcpy(r, t, sizeof t);
}
inline void q2m(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
inline void rot(double* r, const double* q, const double* v) {
  double m[9]; q2m(m, q);
  double t[3] = {m[0] * v[0] + m[1] * v[1] + m[2] * v[2], m[3] * v[0] + m[4] * v[1] + m[5] * v[2], m[6] * v[0] + m[7] * v[1] + m[8] * v[2]};
  std::memcpy(r, t, sizeof t);
}
// (pa,qa) o (pb,qb)
inline void compose(double* p, double* q, const double* pa, const double* qa, const double* pb, const double* qb) {
  double v[3]; rot(v, qa, pb);
  double pp[3] = {pa[0] + v[0], pa[1] + v[1], pa[2] + v[2]}, qq[4]; qmul(qq, qa, qb);
  double n = std::sqrt(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
  for (int k = 0; k < 4; ++k) q[k] = qq[k] / n;
  std::memcpy(p, pp, sizeof pp);
}
// symmetric 3x3 eigen-decomposition (cyclic Jacobi); V columns = eigenvectors
inline void eig3(const double* A, double* w, double* V) {
  double a[9]; std::memcpy(a, A, sizeof a);
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 50; ++sweep) {
    double off = std::fabs(a[1]) + std::fabs(a[2]) + std::fabs(a[5]);
    if (off < 1e-30) break;
    for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 3; ++q) {
      double apq = a[3 * p + q]; if (std::fabs(apq) < 1e-300) continue;
      double th = (a[3 * q + q] - a[3 * p + p]) / (2 * apq), t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < 3; ++k) { double akp = a[3 * k + p], akq = a[3 * k + q]; a[3 * k + p] = c * akp - s * akq; a[3 * k + q] = s * akp + c * akq; }
      for (int k = 0; k < 3; ++k) { double apk = a[3 * p + k], aqk = a[3 * q + k]; a[3 * p + k] = c * apk - s * aqk; a[3 * q + k] = s * apk + c * aqk; }
      for (int k = 0; k < 3; ++k) { double vkp = V[3 * k + p], vkq = V[3 * k + q]; V[3 * k + p] = c * vkp - s * vkq; V[3 * k + q] = s * vkp + c * vkq; }
    }
  }
  w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
}
inline void m2q(double* q, const double* R) {
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) { double s = std::sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); for (int k = 0; k < 4; ++k) q[k] /= n;
}
}  // namespace merge_detail

// returns the merged model; body_map[old body] = new body
inline HostModel merge_fixed_bodies(const HostModel& h, std::vector<int>& body_map) {
  using namespace merge_detail;
  const int nb = h.nbody;
  const auto &par = h.I("body_parentid"), &bjn = h.I("body_jntnum"), &weld = h.I("body_weldid");
  const auto &bpos = h.D("body_pos"), &bquat = h.D("body_quat"), &ipos = h.D("body_ipos"), &iquat = h.D("body_iquat"), &mass = h.D("body_mass"), &inertia = h.D("body_inertia");
  std::vector<int> keep(nb), anc(nb), newid(nb, -1);
  std::vector<double> rp(3 * nb, 0.0), rq(4 * nb, 0.0);
  int nn = 0;
  for (int b = 0; b < nb; ++b) {
    keep[b] = b == 0 || bjn[b] > 0 || weld[b] == 0;
    if (keep[b]) { anc[b] = b; rq[4 * b] = 1; newid[b] = nn++; }
    else { int p = par[b]; anc[b] = anc[p]; compose(&rp[3 * b], &rq[4 * b], &rp[3 * p], &rq[4 * p], &bpos[3 * b], &bquat[4 * b]); }
  }
  body_map.assign(nb, 0);
  for (int b = 0; b < nb; ++b) body_map[b] = newid[anc[b]];
  if (nn == nb) return h;
  HostModel o = h;
  o.nbody = nn;
  std::vector<int> nparent(nn), njadr(nn, -1), njnum(nn, 0), ndadr(nn, -1), ndnum(nn, 0), nroot(nn, 0), nweld(nn, 0);
  std::vector<double> npos(3 * nn), nquat(4 * nn), nipos(3 * nn, 0.0), niquat(4 * nn, 0.0), nmass(nn, 0.0), ninertia(3 * nn, 0.0), ninvw(2 * nn, 0.0);
  std::vector<std::string> nnames(nn);
  for (int b = 0; b < nb; ++b) if (keep[b]) {
    int k = newid[b], p = par[b];
    nparent[k] = b == 0 ? 0 : body_map[p];
    if (b == 0) { for (int i = 0; i < 3; ++i) npos[i] = 0; nquat[0] = 1; nquat[1] = nquat[2] = nquat[3] = 0; }
    else compose(&npos[3 * k], &nquat[4 * k], &rp[3 * p], &rq[4 * p], &bpos[3 * b], &bquat[4 * b]);
    njadr[k] = h.I("body_jntadr")[b]; njnum[k] = bjn[b]; ndadr[k] = h.I("body_dofadr")[b]; ndnum[k] = h.I("body_dofnum")[b];
    nnames[k] = h.names.at(OBJ_BODY)[b];
    ninvw[2 * k] = h.D("body_invweight0")[2 * b]; ninvw[2 * k + 1] = h.D("body_invweight0")[2 * b + 1];
    // composite mass properties of the members
    std::vector<int> mem; for (int c = b; c < nb; ++c) if (anc[c] == b) mem.push_back(c);
    if (mem.size() == 1) {
      nmass[k] = mass[b]; for (int i = 0; i < 3; ++i) { nipos[3 * k + i] = ipos[3 * b + i]; ninertia[3 * k + i] = inertia[3 * b + i]; }
      for (int i = 0; i < 4; ++i) niquat[4 * k + i] = iquat[4 * b + i];
    } else {
      double M = 0, com[3] = {0, 0, 0};
      std::vector<double> cpos(3 * mem.size());
      for (size_t j = 0; j < mem.size(); ++j) {
        int c = mem[j]; double v[3]; rot(v, &rq[4 * c], &ipos[3 * c]);
        for (int i = 0; i < 3; ++i) { cpos[3 * j + i] = rp[3 * c + i] + v[i]; com[i] += mass[c] * cpos[3 * j + i]; }
        M += mass[c];
      }
      if (M > 0) for (int i = 0; i < 3; ++i) com[i] /= M;
      double I[9] = {0};
      for (size_t j = 0; j < mem.size(); ++j) {
        int c = mem[j]; double q[4], R[9]; qmul(q, &rq[4 * c], &iquat[4 * c]); q2m(R, q);
        double d[3] = {cpos[3 * j] - com[0], cpos[3 * j + 1] - com[1], cpos[3 * j + 2] - com[2]}, dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) {
          double v = 0; for (int t = 0; t < 3; ++t) v += R[3 * r + t] * inertia[3 * c + t] * R[3 * s + t];
          I[3 * r + s] += v + mass[c] * ((r == s ? dd : 0.0) - d[r] * d[s]);
        }
      }
      double w[3], V[9]; eig3(I, w, V);
      double det = V[0] * (V[4] * V[8] - V[5] * V[7]) - V[1] * (V[3] * V[8] - V[5] * V[6]) + V[2] * (V[3] * V[7] - V[4] * V[6]);
      if (det < 0) { V[2] = -V[2]; V[5] = -V[5]; V[8] = -V[8]; }
      nmass[k] = M; for (int i = 0; i < 3; ++i) { nipos[3 * k + i] = com[i]; ninertia[3 * k + i] = w[i]; }
      m2q(&niquat[4 * k], V);
    }
  }
  for (int k = 1; k < nn; ++k) { int p = nparent[k]; nroot[k] = p == 0 ? k : nroot[p]; nweld[k] = njnum[k] > 0 ? k : nweld[p]; }
  auto setd = [&](const char* n, const std::vector<double>& v, std::vector<long long> sh) { auto& a = o.arr[n]; a.d = v; a.shape = sh; a.is_int = false; };
  auto seti = [&](const char* n, const std::vector<int>& v, std::vector<long long> sh) { auto& a = o.arr[n]; a.i = v; a.shape = sh; a.is_int = true; };
  seti("body_parentid", nparent, {nn}); seti("body_jntadr", njadr, {nn}); seti("body_jntnum", njnum, {nn}); seti("body_dofadr", ndadr, {nn}); seti("body_dofnum", ndnum, {nn});
  seti("body_rootid", nroot, {nn}); seti("body_weldid", nweld, {nn});
  setd("body_pos", npos, {nn, 3}); setd("body_quat", nquat, {nn, 4}); setd("body_ipos", nipos, {nn, 3}); setd("body_iquat", niquat, {nn, 4});
  setd("body_mass", nmass, {nn}); setd("body_inertia", ninertia, {nn, 3}); setd("body_invweight0", ninvw, {nn, 2});
  o.names[OBJ_BODY] = nnames;
  for (auto& v : o.I("jnt_bodyid")) v = body_map[v];
  for (auto& v : o.I("dof_bodyid")) v = body_map[v];
  auto reattach = [&](const char* bodyid, const char* pos, const char* quat, int n) {
    auto& bid = o.I(bodyid); auto& P = o.D(pos); auto& Q = o.D(quat);
    for (int g = 0; g < n; ++g) { int b = h.I(bodyid)[g]; compose(&P[3 * g], &Q[4 * g], &rp[3 * b], &rq[4 * b], &h.D(pos)[3 * g], &h.D(quat)[4 * g]); bid[g] = body_map[b]; }
  };
  reattach("geom_bodyid", "geom_pos", "geom_quat", h.ngeom);
  reattach("site_bodyid", "site_pos", "site_quat", h.nsite);
  for (int e = 0; e < h.neq; ++e) if (h.I("eq_type")[e] == EQ_CONNECT) {
    for (int side = 0; side < 2; ++side) {
      int b = side == 0 ? h.I("eq_obj1id")[e] : h.I("eq_obj2id")[e];
      double v[3]; rot(v, &rq[4 * b], &h.D("eq_data")[11 * e + 3 * side]);
      for (int i = 0; i < 3; ++i) o.D("eq_data")[11 * e + 3 * side + i] = rp[3 * b + i] + v[i];
      (side == 0 ? o.I("eq_obj1id") : o.I("eq_obj2id"))[e] = body_map[b];
    }
  }
  return o;
}

}  // namespace ur3e
