"""Oracle-backed stand-in for ecb200.Engine (TEST INFRASTRUCTURE for the GPU-less container only).

The full-size GPU tests and the workload generators are Python around the engine's batch methods; this class answers
the same methods from oracle/ecoracle.py so that their logic (edge rows, expected masks, samplers, libcrypto
comparisons) can be exercised at tiny sizes on the CPU tier before GPU time is spent.  It is never imported by the
product, by bench.py or by any -m gpu test."""
from oracle import ecoracle as o

FLAG_CT, FLAG_COMPRESSED, FLAG_UNCOMPRESSED, FLAG_PROJ = 1, 2, 4, 8


def _compress(c, flags):
    return True if flags & FLAG_COMPRESSED else False if flags & FLAG_UNCOMPRESSED else None


def _b(x):
    return x.tobytes() if hasattr(x, "tobytes") else bytes(x)


class OracleEngine:
    n_devices = 1

    def mul_by_generator_batch(self, curve, ks, flags=0):
        c = o.curve(curve)
        return o.batch_mul_gen(c, _b(ks), _compress(c, flags))

    def mul_batch(self, curve, points, ks, inf=None, flags=0):
        c = o.curve(curve)
        ks, points = _b(ks), _b(points)
        n = len(ks) // c.fb
        if flags & FLAG_PROJ:
            return o.batch_mul_var_proj(c, points, ks, _compress(c, flags)), bytes(n)
        return o.batch_mul_var_affine(c, points, None if inf is None else _b(inf), ks, _compress(c, flags)), bytes(n)

    def batch_normalize(self, curve, xyz):
        c = o.curve(curve)
        xyz = _b(xyz)
        fb = c.fb
        n = len(xyz) // (3 * fb)
        pts = [tuple(int.from_bytes(xyz[(3 * i + j) * fb:(3 * i + j + 1) * fb], "big") for j in range(3)) for i in range(n)]
        xy, inf = bytearray(), bytearray()
        for P in o.batch_normalize(c, pts):
            if P is None:
                xy += bytes(2 * fb); inf.append(1)
            else:
                xy += P[0].to_bytes(fb, "big") + P[1].to_bytes(fb, "big"); inf.append(0)
        return bytes(xy), bytes(inf)

    def field_op(self, curve, which, op, a, b=None):
        c = o.curve(curve)
        m = c.n if which == 1 else c.p
        fb = c.fb
        a = _b(a)
        b = a if b is None else _b(b)
        n = len(a) // fb
        out, ok = bytearray(), bytearray()
        f = {0: lambda x, y: (x + y) % m, 1: lambda x, y: (x - y) % m, 2: lambda x, y: x * y % m, 3: lambda x, y: x * x % m,
             4: lambda x, y: (-x) % m, 5: lambda x, y: pow(x, -1, m) if x % m else 0}[op]
        for i in range(n):
            x, y = int.from_bytes(a[i * fb:(i + 1) * fb], "big"), int.from_bytes(b[i * fb:(i + 1) * fb], "big")
            # the device reduces once on entry for the scalar-field "add zero" idiom of the generators; mirror that leniency
            out += f(x % m, y % m).to_bytes(fb, "big")
            ok.append(1 if x < m and y < m else 0)
        return bytes(out), bytes(ok)

    def ecdsa_verify(self, curve, q, z, rs, out=None):
        c = o.curve(curve)
        ok = o.batch_verify(c, _b(q), _b(z), _b(rs))
        if out is not None:
            out[:] = memoryview(ok)
            return out
        return ok

    def lincomb2_batch(self, curve, p1, k1, p2, k2, flags=0):
        c = o.curve(curve)
        fb = c.fb
        p1, k1, p2, k2 = _b(p1), _b(k1), _b(p2), _b(k2)
        n = len(k1) // fb
        out = bytearray()

        def pt(buf, i):
            if flags & FLAG_PROJ:
                X, Y, Z = (int.from_bytes(buf[(3 * i + j) * fb:(3 * i + j + 1) * fb], "big") for j in range(3))
                return o.proj_to_affine(c, X, Y, Z)
            return (int.from_bytes(buf[2 * i * fb:(2 * i + 1) * fb], "big"), int.from_bytes(buf[(2 * i + 1) * fb:(2 * i + 2) * fb], "big"))

        for i in range(n):
            a = o.reduce_once(c, int.from_bytes(k1[i * fb:(i + 1) * fb], "big"))
            b = o.reduce_once(c, int.from_bytes(k2[i * fb:(i + 1) * fb], "big"))
            out += o.slot_encode(c, o.pt_lincomb(c, [(pt(p1, i), a), (pt(p2, i), b)]), _compress(c, flags))
        return bytes(out), bytes(n)

    def close(self):
        pass
