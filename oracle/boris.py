"""Oracle restatement of ``BorisPush.push`` -- TEST INFRASTRUCTURE ONLY.

Follows turbopy/computetools.py:427-441 operation by operation.  Every line
below is one correctly rounded fp64 operation sequence in the same
association order as the reference's numpy expression (numpy never contracts
a*b+c into an FMA), so both functions are bit-identical to the reference;
``make_golden.py`` asserts that against the imported reference.

Two forms:

``push_aos``  whole-array (N,3) numpy, the way the reference evaluates it
              (two ``np.cross`` calls, axis=-1 sums).  This is the form that is
              *timed* as the CPU baseline, because it has the reference's
              memory behaviour (~25 full-array temporaries).
``push_soa``  component-wise on six length-N arrays: the line-by-line spec of
              the CUDA kernel (``turbopy_b200/csrc/boris_math.cuh``).
"""
import numpy as np

#: turbopy/computetools.py:405 -- note: 2.9979e8, not CODATA c
C2 = 2.9979e8 ** 2


def push_aos(position, momentum, charge, mass, E, B, dt, c2=C2):
    """(N,3) arrays, in place.  computetools.py:427-441."""
    # :429  vminus = momentum + dt * E * charge / 2
    half_kick = ((dt * E) * charge) / 2
    vminus = momentum + half_kick
    # :430-431  gamma*m from the OLD momentum
    m1 = np.sqrt(mass ** 2 + np.sum(momentum * momentum, axis=-1) / c2)
    # :433-434
    t = (((dt * B) * charge) / m1[:, np.newaxis]) / 2
    s = (2 * t) / (1 + np.sum(t * t, axis=-1)[:, np.newaxis])
    # :436-438
    vprime = vminus + np.cross(vminus, t)
    vplus = vminus + np.cross(vprime, s)
    momentum[:] = vplus + half_kick
    # :439-441
    m2 = np.sqrt(mass ** 2 + np.sum(momentum * momentum, axis=-1) / c2)
    position[:] = position + (dt * momentum) / m2[:, np.newaxis]


def push_soa(x, y, z, px, py, pz, charge, mass, Ex, Ey, Ez, Bx, By, Bz,
             dt, c2=C2):
    """Six length-N arrays updated in place; field components are scalars or
    length-N arrays.  Same rounding sequence as ``push_aos``:

    * ``np.sum(v*v, axis=-1)`` over a length-3 axis adds left to right:
      ``(v0*v0 + v1*v1) + v2*v2``.
    * ``np.cross(a, b)`` = ``(a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)``,
      each product rounded on its own.
    """
    hx = ((dt * Ex) * charge) / 2
    hy = ((dt * Ey) * charge) / 2
    hz = ((dt * Ez) * charge) / 2
    vmx = px + hx
    vmy = py + hy
    vmz = pz + hz
    mm = mass ** 2
    m1 = np.sqrt(mm + ((px * px + py * py) + pz * pz) / c2)
    tx = (((dt * Bx) * charge) / m1) / 2
    ty = (((dt * By) * charge) / m1) / 2
    tz = (((dt * Bz) * charge) / m1) / 2
    den = 1 + ((tx * tx + ty * ty) + tz * tz)
    sx = (2 * tx) / den
    sy = (2 * ty) / den
    sz = (2 * tz) / den
    vpx = vmx + (vmy * tz - vmz * ty)
    vpy = vmy + (vmz * tx - vmx * tz)
    vpz = vmz + (vmx * ty - vmy * tx)
    plx = vmx + (vpy * sz - vpz * sy)
    ply = vmy + (vpz * sx - vpx * sz)
    plz = vmz + (vpx * sy - vpy * sx)
    npx = plx + hx
    npy = ply + hy
    npz = plz + hz
    m2 = np.sqrt(mm + ((npx * npx + npy * npy) + npz * npz) / c2)
    px[...] = npx
    py[...] = npy
    pz[...] = npz
    x[...] = x + (dt * npx) / m2
    y[...] = y + (dt * npy) / m2
    z[...] = z + (dt * npz) / m2
