// on-device field update of the particle_in_field example and a
// device-resident clock, so that a whole fundamental_cycle (field update ->
// gather + push -> clock.advance, turbopy/core.py:145-158) can be captured in
// one CUDA graph and replayed K times without the host.
//
//   EMWave.update   tests/fixtures/particle_in_field/particle_in_field.py:32-34
//       phase = -omega*t + k*(r - x0);  E = E0 * cos((2*pi) * phase)
//   SimulationClock.advance   turbopy/core.py:533-538
//       this_step += 1;  time = start_time + dt*this_step
//
// Same operation order as the numpy expressions (-fmad=false).  cos() is the
// CUDA fp64 library function (<= 2 ulp); numpy's cos is a different
// implementation, so the field agrees with the reference to a few ulp, not bit
// for bit (tolerance in tests/test_gpu_cycle.py).
#include "common.cuh"

namespace tp {

__global__ void plane_wave_kernel(const double* __restrict__ r, int64_t ng, double amplitude, double omega, double k,
                                  double x0, const double* __restrict__ clock, double* __restrict__ E,
                                  int64_t stride) {
    const double t = clock[0];
    const double two_pi = 2 * 3.141592653589793;           // 2 * np.pi
    const double wt = -omega * t;
    const int64_t gstride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < ng; j += gstride) {
        const double phase = wt + k * (r[j] - x0);
        E[j * stride] = amplitude * cos(two_pi * phase);
    }
}

// clock[0] = time, clock[1] = this_step (kept as a double: exact up to 2^53)
__global__ void clock_advance_kernel(double* clock, double start, double dt) {
    const double step = clock[1] + 1.0;
    clock[1] = step;
    clock[0] = start + dt * step;
}

}  // namespace tp

using namespace tp;

extern "C" int tp_plane_wave_f64(const double* r, int64_t ng, double amplitude, double omega, double k, double x0,
                                 const double* clock_dev, double* E, int64_t e_stride, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!r || !clock_dev || !E || ng < 1 || e_stride < 1) return TP_E_ARG;
    const int threads = 256;
    int64_t blocks = (ng + threads - 1) / threads;
    if (blocks > static_cast<int64_t>(d.sm_count) * 8) blocks = static_cast<int64_t>(d.sm_count) * 8;
    plane_wave_kernel<<<static_cast<int>(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        r, ng, amplitude, omega, k, x0, clock_dev, E, e_stride);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}

extern "C" int tp_clock_advance_f64(double* clock_dev, double start_time, double dt, tp_stream_t stream) {
    const DeviceProps& d = device_props();
    if (d.status != TP_OK) return d.status;
    if (!clock_dev) return TP_E_ARG;
    clock_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(clock_dev, start_time, dt);
    count_launch();
    return static_cast<int>(cudaGetLastError());
}
