// Host-side flat model produced by the C++ MJCF loader (array names follow MuJoCo's mjModel,
// which is what the reference's Python reads: utils/utils.py, controller/controller_func.py).
#pragma once
#include <map>
#include <string>
#include <vector>

namespace ur3e {

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { EQ_CONNECT = 0, EQ_WELD = 1, EQ_JOINT = 2 };
enum { TRN_JOINT = 0, TRN_TENDON = 3 };
// object kinds for name lookups (values follow mujoco.mjtObj where the reference uses them, utils/utils.py:29-66)
enum { OBJ_BODY = 1, OBJ_JOINT = 3, OBJ_GEOM = 5, OBJ_SITE = 6, OBJ_TENDON = 18, OBJ_ACTUATOR = 19, OBJ_KEY = 23 };

struct HostArray {
  std::vector<double> d;    // float64 payload (or)
  std::vector<int> i;       // int32 payload
  std::vector<long long> shape;
  bool is_int = false;
};

struct HostModel {
  int nq = 0, nv = 0, nu = 0, nbody = 0, njnt = 0, ngeom = 0, nsite = 0, neq = 0, ntendon = 0, nwrap = 0, npair = 0, nkey = 0;
  double timestep = 0.002, gravity[3] = {0, 0, -9.81}, impratio = 1, meaninertia = 1;
  int cone_elliptic = 0;
  std::map<std::string, HostArray> arr;
  std::map<int, std::vector<std::string>> names;  // by OBJ_* kind
  std::vector<std::string> warnings;

  std::vector<double>& D(const std::string& n) { return arr[n].d; }
  std::vector<int>& I(const std::string& n) { arr[n].is_int = true; return arr[n].i; }
  const std::vector<double>& D(const std::string& n) const { return arr.at(n).d; }
  const std::vector<int>& I(const std::string& n) const { return arr.at(n).i; }
  int name2id(int kind, const std::string& n) const {
    auto it = names.find(kind);
    if (it == names.end()) return -1;
    for (size_t k = 0; k < it->second.size(); ++k) if (it->second[k] == n) return (int)k;
    return -1;
  }
};

// Parses an MJCF file (subset used by the reference's scenes) and runs the
// model-constant pass (connect anchors, invweight0, meaninertia).  Throws std::runtime_error.
HostModel load_mjcf(const std::string& path);

}  // namespace ur3e
