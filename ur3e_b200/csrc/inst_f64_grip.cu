#include "batch_impl.cuh"
namespace ur3e {
std::unique_ptr<BatchBase> make_batch_f64_grip(const HostModel& h, const ur3e_env_config& c, long long n, int dev) { return make_batch<double, DimsGrip>(h, c, n, dev); }
}  // namespace ur3e
