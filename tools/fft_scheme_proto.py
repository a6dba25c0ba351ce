"""Index-level prototype of the 16-lanes-per-frame real-FFT scheme used by
telugu_asr_b200/csrc/logmel.cu (development aid; checks the decomposition against
numpy's rfft).  Not used by the product or the tests."""
import numpy as np

def fft4(p0, p1, p2, p3):
    s0, s1, s2, s3 = p0 + p2, p0 - p2, p1 + p3, p1 - p3
    return s0 + s2, s1 - 1j * s3, s0 - s2, s1 + 1j * s3

def fft16(v):
    v = list(v)
    for a in range(4):
        v[a], v[a + 4], v[a + 8], v[a + 12] = fft4(v[a], v[a + 4], v[a + 8], v[a + 12])
    for a in range(4):
        for d in range(4):
            v[a + 4 * d] *= np.exp(-2j * np.pi * a * d / 16)
    for d in range(4):
        v[4 * d], v[1 + 4 * d], v[2 + 4 * d], v[3 + 4 * d] = fft4(v[4 * d], v[1 + 4 * d], v[2 + 4 * d], v[3 + 4 * d])
    # X[4c+d] sits at v[c+4d]
    return [v[(k >> 2) + 4 * (k & 3)] for k in range(16)]

rng = np.random.default_rng(0)
x = rng.standard_normal(16) + 1j * rng.standard_normal(16)
assert np.allclose(fft16(x), np.fft.fft(x))

frame = rng.standard_normal(400)
hw = np.zeros(512); hw[:400] = 0.5 * (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400))
y = np.zeros(512); y[:400] = frame
ref = np.abs(np.fft.rfft(y * 2 * hw)) ** 2

# pass 1: lane t owns m1=t, inputs z[t+16*m2]
A = np.zeros((16, 16), complex)      # [t][k2]
for t in range(16):
    v = [(y[2 * (t + 16 * m2)] * hw[2 * (t + 16 * m2)] + 1j * y[2 * (t + 16 * m2) + 1] * hw[2 * (t + 16 * m2) + 1]) if m2 < 13 else 0j for m2 in range(16)]
    out = fft16(v)
    for k2 in range(16):
        A[t][k2] = out[k2] * np.exp(-2j * np.pi * ((t * k2) & 255) / 256)
# transpose through scratch[k2*17 + t]
scratch = np.zeros(16 * 17, complex)
for t in range(16):
    for k2 in range(16):
        scratch[k2 * 17 + t] = A[t][k2]
Z = np.zeros((16, 16), complex)      # [t][k1] = Z[t+16k1]
for t in range(16):
    Z[t] = fft16([scratch[t * 17 + n1] for n1 in range(16)])
# check Z against the complex FFT of the packed sequence (half-scaled)
zz = (y * hw)[0::2] + 1j * (y * hw)[1::2]
ZZ = np.fft.fft(zz)
for t in range(16):
    for k1 in range(16):
        assert np.allclose(Z[t][k1], ZZ[t + 16 * k1])
# pairing
P = np.full(257, np.nan)
for t in range(16):
    partner = (16 - t) & 15
    base = np.exp(-2j * np.pi * t / 512)
    w0 = -1j if t == 0 else base
    for j in range(8):
        a = Z[t][j]
        b = Z[partner][15 - j]
        if t == 0:
            if j == 0:
                a = Z[0][8]; b = Z[0][8]
            else:
                b = Z[0][16 - j]
        E = a + np.conj(b)
        O = -1j * (a - np.conj(b))
        W = w0 if j == 0 else base * np.exp(-2j * np.pi * j / 32)
        T = W * O
        ka = 128 if (t == 0 and j == 0) else t + 16 * j
        kb = 256 - ka
        P[ka] = abs(E + T) ** 2
        P[kb] = abs(E - T) ** 2
    if t == 0:
        z0 = Z[0][0]
        P[0] = 4 * (z0.real + z0.imag) ** 2
        P[256] = 4 * (z0.real - z0.imag) ** 2
assert not np.isnan(P).any()
print("max rel err", np.max(np.abs(P - ref) / (np.abs(ref) + 1e-30)))
assert np.allclose(P, ref, rtol=1e-9, atol=1e-9)
print("scheme OK")
