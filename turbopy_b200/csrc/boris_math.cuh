// BorisPush.push arithmetic for one particle, in the reference's operation
// order (turbopy/computetools.py:427-441; line-by-line spec: oracle/boris.py
// push_soa).  Compiled with -fmad=false: every * and + below rounds on its own.
#pragma once
#include "common.cuh"

namespace tp {

struct PushK {            // kernel-side constants of one push call
    double dt, q, mm;     // clock.dt, charge, mass**2
    double c2, rc2;       // BorisPush.c2 and RN(1/c2) (host-rounded)
    int    fast_ok;       // c2 is usable by the FAST division policy
};

struct Vec3 { double x, y, z; };

// (dt*F)*q : the field-dependent numerator shared by the half kick (E) and the
// rotation vector (B); computetools.py:429,433 evaluate dt*F first, then *charge.
__device__ __forceinline__ Vec3 dt_field_q(const PushK& k, const Vec3& F) {
    Vec3 r;
    r.x = (k.dt * F.x) * k.q;
    r.y = (k.dt * F.y) * k.q;
    r.z = (k.dt * F.z) * k.q;
    return r;
}

// One Boris step.  eq = (dt*E)*q, bq = (dt*B)*q (see dt_field_q); x and p are
// updated in place.  Divisions by 2 are exact multiplications by 0.5.  FAST:
// the quotients that share a denominator (m1, 1+|t|^2, m2, c2) use one
// correctly rounded reciprocal each and record their validity in `g`
// (common.cuh); the caller must test guard_ok(g) && finite3(x) and recompute
// with FAST = false otherwise.  FAST = false is plain IEEE arithmetic.
template <bool FAST>
__device__ __forceinline__ void boris_step(const PushK& k, const Vec3& eq, const Vec3& bq,
                                           Vec3& x, Vec3& p, Guard& g) {
    const Recip<FAST> rc2 = const_recip<FAST>(k.c2, k.rc2);
    // :429  vminus = momentum + dt*E*charge/2
    const double hx = eq.x * 0.5, hy = eq.y * 0.5, hz = eq.z * 0.5;
    const double vmx = p.x + hx, vmy = p.y + hy, vmz = p.z + hz;
    // :430-431  m1 = sqrt(mass**2 + sum(p*p)/c2)   (old momentum)
    const double pp = (p.x * p.x + p.y * p.y) + p.z * p.z;
    const double m1 = root<FAST>(k.mm + quot<FAST>(pp, rc2, g), g);
    // :433  t = dt*B*charge/m1/2
    const Recip<FAST> r1 = make_recip<FAST>(m1, g);
    const double ux = quot<FAST>(bq.x, r1, g), uy = quot<FAST>(bq.y, r1, g), uz = quot<FAST>(bq.z, r1, g);
    const double tx = ux * 0.5, ty = uy * 0.5, tz = uz * 0.5;
    // :434  s = 2*t/(1 + sum(t*t))
    const double den = 1.0 + ((tx * tx + ty * ty) + tz * tz);
    const Recip<FAST> rd = make_recip<FAST>(den, g);
    // 2*t: on the FAST path it is u itself -- halving and doubling are exact unless u*0.5 lost a bit to underflow
    // (|u| < 2^-1021), and such a numerator trips the guard either way; the IEEE path forms 2*t literally.
    const double sx = quot<FAST>(FAST ? ux : 2.0 * tx, rd, g);
    const double sy = quot<FAST>(FAST ? uy : 2.0 * ty, rd, g);
    const double sz = quot<FAST>(FAST ? uz : 2.0 * tz, rd, g);
    // :436  vprime = vminus + cross(vminus, t)
    const double vpx = vmx + (vmy * tz - vmz * ty);
    const double vpy = vmy + (vmz * tx - vmx * tz);
    const double vpz = vmz + (vmx * ty - vmy * tx);
    // :437  vplus = vminus + cross(vprime, s)
    const double plx = vmx + (vpy * sz - vpz * sy);
    const double ply = vmy + (vpz * sx - vpx * sz);
    const double plz = vmz + (vpx * sy - vpy * sx);
    // :438  momentum = vplus + dt*E*charge/2
    p.x = plx + hx;
    p.y = ply + hy;
    p.z = plz + hz;
    // :439-440  m2 from the new momentum
    const double qq = (p.x * p.x + p.y * p.y) + p.z * p.z;
    const double m2 = root<FAST>(k.mm + quot<FAST>(qq, rc2, g), g);
    // :441  position = position + dt*momentum/m2
    const Recip<FAST> r2 = make_recip<FAST>(m2, g);
    x.x = x.x + quot<FAST>(k.dt * p.x, r2, g);
    x.y = x.y + quot<FAST>(k.dt * p.y, r2, g);
    x.z = x.z + quot<FAST>(k.dt * p.z, r2, g);
}

}  // namespace tp
