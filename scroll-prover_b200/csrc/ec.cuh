// BN254 G1 (y^2 = x^3 + 3 over Fq) on the device.
//
// Replaces halo2curves::bn256::{G1Affine, G1} arithmetic (halo2curves 0.1.0 @ 112f5b9, pin
// /root/reference/Cargo.lock:1911-1913; src/bn256/curve.rs, src/derive/curve.rs).  Layouts at the ABI
// are the Rust ones: affine (x, y) 64 B with identity (0,0); Jacobian (x, y, z) 96 B with z = 0 identity.
// Internally bucket sums use extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// mixed add 8M+2S, full add 12M+2S, identity encoded as ZZ = 0.  Group elements are
// representation-independent, so results are compared (and returned) only after normalisation.
#pragma once
#include "ff.cuh"

namespace b200zk {

struct Affine {
    Fq x, y;
    FF_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct XYZZ {
    Fq x, y, zz, zzz;
    FF_HD static XYZZ identity() {
        XYZZ r;
        r.x = Fq::zero();
        r.y = Fq::zero();
        r.zz = Fq::zero();
        r.zzz = Fq::zero();
        return r;
    }
    FF_HD bool is_identity() const { return zz.is_zero(); }
};

struct Jacobian {
    Fq x, y, z;
};

// 2 * (affine p), p != identity   (mdbl-2008-s-1)
FF_HD XYZZ xyzz_dbl_affine(const Fq& px, const Fq& py) {
    XYZZ r;
    Fq U = py.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = px * V;
    Fq X2 = px.sqr();
    Fq M = X2.dbl() + X2;
    r.x = M.sqr() - S.dbl();
    r.y = M * (S - r.x) - W * py;
    r.zz = V;
    r.zzz = W;
    return r;
}

// 2 * p   (dbl-2008-s-1, a = 0)
FF_HD XYZZ xyzz_dbl(const XYZZ& p) {
    if (p.is_identity()) return p;
    XYZZ r;
    Fq U = p.y.dbl();
    Fq V = U.sqr();
    Fq W = U * V;
    Fq S = p.x * V;
    Fq X2 = p.x.sqr();
    Fq M = X2.dbl() + X2;
    r.x = M.sqr() - S.dbl();
    r.y = M * (S - r.x) - W * p.y;
    r.zz = V * p.zz;
    r.zzz = W * p.zzz;
    return r;
}

// acc += (qx, qy) affine, q != identity   (madd-2008-s with the exceptional cases)
FF_HD void xyzz_madd(XYZZ& acc, const Fq& qx, const Fq& qy) {
    if (acc.is_identity()) {
        acc.x = qx;
        acc.y = qy;
        acc.zz = Fq::one();
        acc.zzz = Fq::one();
        return;
    }
    Fq U2 = qx * acc.zz;
    Fq S2 = qy * acc.zzz;
    Fq P = U2 - acc.x;
    Fq R = S2 - acc.y;
    if (P.is_zero()) {
        if (R.is_zero())
            acc = xyzz_dbl_affine(qx, qy);
        else
            acc = XYZZ::identity();
        return;
    }
    Fq PP = P.sqr();
    Fq PPP = P * PP;
    Fq Q = acc.x * PP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
    Fq Y3 = R * (Q - X3) - acc.y * PPP;
    acc.x = X3;
    acc.y = Y3;
    acc.zz = acc.zz * PP;
    acc.zzz = acc.zzz * PPP;
}

// acc += q   (add-2008-s with the exceptional cases)
FF_HD void xyzz_add(XYZZ& acc, const XYZZ& q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) {
        acc = q;
        return;
    }
    Fq U1 = acc.x * q.zz;
    Fq U2 = q.x * acc.zz;
    Fq S1 = acc.y * q.zzz;
    Fq S2 = q.y * acc.zzz;
    Fq P = U2 - U1;
    Fq R = S2 - S1;
    if (P.is_zero()) {
        if (R.is_zero())
            acc = xyzz_dbl(acc);
        else
            acc = XYZZ::identity();
        return;
    }
    Fq PP = P.sqr();
    Fq PPP = P * PP;
    Fq Q = U1 * PP;
    Fq X3 = R.sqr() - PPP - Q.dbl();
    Fq Y3 = R * (Q - X3) - S1 * PPP;
    acc.x = X3;
    acc.y = Y3;
    acc.zz = acc.zz * q.zz * PP;
    acc.zzz = acc.zzz * q.zzz * PPP;
}

FF_HD XYZZ xyzz_from_affine(const Affine& a) {
    XYZZ r = XYZZ::identity();
    if (!a.is_identity()) {
        r.x = a.x;
        r.y = a.y;
        r.zz = Fq::one();
        r.zzz = Fq::one();
    }
    return r;
}

// Jacobian (X, Y, Z) -> XYZZ: ZZ = Z^2, ZZZ = Z^3 (same X, Y)
FF_HD XYZZ xyzz_from_jacobian(const Jacobian& j) {
    XYZZ r = XYZZ::identity();
    if (!j.z.is_zero()) {
        r.x = j.x;
        r.y = j.y;
        r.zz = j.z.sqr();
        r.zzz = r.zz * j.z;
    }
    return r;
}

FF_HD Affine xyzz_to_affine(const XYZZ& p) {
    Affine a;
    if (p.is_identity()) {
        a.x = Fq::zero();
        a.y = Fq::zero();
        return a;
    }
    Fq i = (p.zz * p.zzz).inv();
    a.x = p.x * (i * p.zzz);  // X / ZZ
    a.y = p.y * (i * p.zz);   // Y / ZZZ
    return a;
}

// normalised Jacobian: (x, y, 1), identity = (0, 1, 0) as halo2curves' G1::identity()
FF_HD Jacobian xyzz_to_jacobian_normalized(const XYZZ& p) {
    Jacobian j;
    if (p.is_identity()) {
        j.x = Fq::zero();
        j.y = Fq::one();
        j.z = Fq::zero();
        return j;
    }
    Affine a = xyzz_to_affine(p);
    j.x = a.x;
    j.y = a.y;
    j.z = Fq::one();
    return j;
}

}  // namespace b200zk
