#include "batch_impl.cuh"
namespace ur3e {
std::unique_ptr<BatchBase> make_batch_f32_main(const HostModel& h, const ur3e_env_config& c, long long n, int dev) { return make_batch<float, DimsMain, DimsMainLite, DimsMainMid>(h, c, n, dev); }
}  // namespace ur3e
