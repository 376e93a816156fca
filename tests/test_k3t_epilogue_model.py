"""CPU model of the epilogue of k3_tma_kernel (wignerd/jl_b200/csrc/wd_k3_tma.cuh: k3t_emit / k3t_box): the lane-level index,
sign and phase arithmetic of the kernel restated in Python and run against the oracle.  What it pins, without a GPU:
  * every element of every matrix is written exactly once by the four boxes of the units (no gaps, no overlaps, also for the
    partial last 16-row group, for m = 0 / n = 0 of integer j and with the row-pair swap of lanes c = 2, 3);
  * the values -- butterfly results times the factorised phase (ca - i sa)(cg -+ i sg), the sign sigma of each target and the
    conjugate relation between a value's two targets -- reproduce wignerD of the reference (src/WignerD.jl:335-342).
The DMMA main loop is not modelled: S and D are taken from the oracle's d (d[i,i'] = sigma(i-i') S, d[i,i'r] = sigma(i-i'r) D)."""
import numpy as np
import pytest

from oracle import wignerd_oracle as o


def sigma_bit(d):
    return (0x6 >> (d & 3)) & 1            # wd_kernels.cuh: sigma_bit (1 = negative)


def flip(v, bit):
    return -v if bit else v


def perm(q):                               # k3t_perm (K3T_SWAP = 1)
    return q ^ ((q >> 3) & 1)


def model(twoj, al, be, ga):
    N = twoj + 1
    Q = (N + 1) // 2
    i0 = N - Q
    zskip = 0 if twoj & 1 else 1
    GC = (Q + 15) >> 4
    qp = 16 * GC + 2
    nb = 8
    dm = o.wignerd_batch(twoj / 2, be)                               # [angle, row, col]
    mem = np.full((nb, 2 * N * N), np.nan)                           # the output as the tensor map sees it
    mhalf = 0.5 * (twoj & 1)
    ph = np.full((4, 8, qp), np.nan)                                 # phase tables of the item (prefetch lambda)
    for which, angs in ((0, al), (1, ga)):
        for ang in range(8):
            for q in range(Q):
                x = (q + mhalf) * angs[ang]
                ph[2 * which, ang, perm(q)] = np.cos(x)
                ph[2 * which + 1, ang, perm(q)] = np.sin(x)

    def sig(i, ip):
        return -1.0 if ((i - ip) & 3) in (1, 2) else 1.0

    for qn in range(Q):                                              # one unit per quadrant column
        ip = i0 + qn
        ipr = N - 1 - ip
        for isd in (False, True):                                    # k3t_emit<GC, ISD>
            boxes = {False: np.full(8 * 2 * Q, np.nan), True: np.full(8 * 2 * Q, np.nan)}
            for lane in range(32):
                r, c = lane >> 2, lane & 3
                nvl = max(0, min(4, Q - 16 * (GC - 1) - 4 * c))
                swb = 1 if (c & 2) else 0
                cg, sg0 = ph[2, r, perm(qn)], ph[3, r, perm(qn)]
                sgx = sg0 if isd else -sg0
                d0 = (qn + 1 - zskip) if isd else -qn
                ms = [sigma_bit(d0 + (jj ^ swb)) for jj in range(4)]
                me = [(d0 + (jj ^ swb)) & 1 for jj in range(4)]
                for g in range(GC):
                    for jj in range(4):
                        q = 16 * g + 4 * c + (jj ^ swb)              # the row that register jj holds
                        if q < Q:
                            i = i0 + q
                            v = sig(i, ipr) * dm[r, i, ipr] if isd else sig(i, ip) * dm[r, i, ip]
                        else:
                            v = np.nan
                        ca, sa = ph[0, r, 16 * g + 4 * c + jj], ph[1, r, 16 * g + 4 * c + jj]
                        y, z = sa * sgx, sa * cg
                        re, im = ca * cg + y, ca * sgx - z
                        vv = flip(v, ms[jj])
                        A, B = vv * re, vv * im
                        for desc in (False, True):                   # k3t_box<GC, DESC>
                            rs = 2 * (i0 if desc else Q)
                            base = r * rs + (2 * (Q - 1 - 4 * c) if desc else 8 * c)
                            base_e = base - 2 * swb if desc else base + 2 * swb
                            base_o = base + 2 * swb if desc else base - 2 * swb
                            skip0 = desc and zskip and c == 0
                            ok = (g < GC - 1 or (jj ^ swb) < nvl) and not (g == 0 and jj == 0 and skip0)
                            bp = base_o if (jj & 1) else base_e
                            if ok:
                                if desc:
                                    off, val = bp - 2 * (16 * g + jj), (flip(A, me[jj]), flip(B, me[jj] ^ 1))
                                else:
                                    off, val = bp + 2 * (16 * g + jj), (A, B)
                                assert 0 <= off and off + 1 < 8 * rs
                                assert np.isnan(boxes[desc][off])    # no lane writes a box element twice
                                boxes[desc][off], boxes[desc][off + 1] = val
            c0, c1, mirror = i0 + qn, Q - 1 - qn, qn >= zskip
            for desc in (False, True):
                do = (isd or mirror) if desc else ((not isd) or mirror)
                if not do:
                    continue
                col = (c0 if isd else c1) if desc else (c1 if isd else c0)
                w = 2 * (i0 if desc else Q)                          # box width of the tensor map (doubles)
                coord0 = 2 * (col * N + (0 if desc else i0))
                for x in range(8):
                    seg = boxes[desc][x * w:(x + 1) * w]
                    assert not np.isnan(seg).any()                   # the box is complete when it is sent off
                    assert np.isnan(mem[x, coord0:coord0 + w]).all() # no element of the output is stored twice
                    mem[x, coord0:coord0 + w] = seg
    assert not np.isnan(mem).any()                                   # every element of every matrix was stored
    return (mem[:, 0::2] + 1j * mem[:, 1::2]).reshape(nb, N, N).transpose(0, 2, 1)


@pytest.mark.parametrize("twoj", [5, 6, 7, 20, 33, 63])
def test_k3t_epilogue_model_reproduces_wignerD(twoj):
    rng = np.random.default_rng(twoj)
    al, be, ga = rng.uniform(0, 6.3, 8), rng.uniform(0, 3.14, 8), rng.uniform(0, 6.3, 8)
    got = model(twoj, al, be, ga)
    ref = o.wignerD_batch(twoj / 2, al, be, ga)
    assert np.abs(got - ref).max() < 1e-13


def test_k3t_epilogue_model_zero_euler_angles_are_exact():
    """alpha = gamma = 0: the phase factors are exactly 1 and 0, so the complex matrix carries the butterfly values unchanged
    (the kernel then equals the real-d kernels bit for bit, test/runtests.jl:343): the quadrant m, n >= 0, which the model takes
    from the oracle, comes back bit for bit, the mirrored quadrants are its signed images"""
    be = np.linspace(0.1, 3.0, 8)
    z = np.zeros(8)
    for twoj in (5, 6):
        got = model(twoj, z, be, z)
        ref = o.wignerd_batch(twoj / 2, be)
        i0 = (twoj + 1) - (twoj + 2) // 2
        assert np.array_equal(got.imag, np.zeros_like(got.imag))
        assert np.array_equal(got.real[:, i0:, i0:], ref[:, i0:, i0:])
        assert np.abs(got.real - ref).max() < 1e-14

