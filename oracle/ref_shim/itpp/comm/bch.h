// itpp/comm/bch.h -- TEST INFRASTRUCTURE.  Stand-in for the slice of IT++ that gr-amps uses: itpp::bin, itpp::bvec
// (string constructor, size, [], (a,b) sub-vector, concat, <<) and itpp::BCH(n, t, systematic) with
// encode(bvec) / decode(bvec, bvec&, bvec&).  IT++ is a third-party dependency that is NOT under /root/reference
// (find_package(ITPP), reference CMakeLists.txt:89; README.md:22 names the distro package libitpp-dev, i.e. 4.3.x) and
// is not installed in this image, so the UNMODIFIED reference sources are compiled against this header instead
// (oracle/Makefile, target _ref).
//
// The BCH class restates IT++ 4.3's published algorithm in IT++'s own representation -- field elements as exponents
// of alpha with -1 for zero, polynomials over GF(2^m) -- so that it is structurally independent of oracle/bch63.c
// (bit-serial LFSR encoder, integer field elements) and of the closed-form syndrome test in the CUDA kernel:
//   * GF(2^m), m = 3..8, on IT++'s primitive polynomials (m = 6: x^6 + x + 1);
//   * g(x) = lcm of the minimal polynomials of alpha^1 .. alpha^(2t)  (n = 63, t = 2: octal 12471, k = 51);
//   * systematic encode: message bits first (first bit = highest power), then x^(n-k) m(x) mod g(x);
//   * decode: S_j = r(alpha^j), j = 1..2t; Berlekamp's simplified iteration run for exactly t steps; Chien search
//     over all n positions; failure when #roots != deg Lambda; the corrected word is re-validated; on failure the
//     systematic part of the received word is handed back.  Returns true iff every block decoded.
// Call sites: lib/focc_impl.cc:105,156-176; lib/fvc_impl.cc:57,98-107; lib/recc_decode_impl.cc:33,53-79.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace itpp {

class bin {
public:
    bin() : b(0) {}
    bin(int v) : b((char)(v & 1)) {}
    operator int() const { return b; }
    bin &operator+=(const bin &o) { b ^= o.b; return *this; }
private:
    char b;
};

class bvec {
public:
    bvec() {}
    explicit bvec(int n) : d(n > 0 ? n : 0) {}
    bvec(const char *s) { parse(s); }
    bvec(const std::string &s) { parse(s.c_str()); }
    int size() const { return (int)d.size(); }
    int length() const { return (int)d.size(); }
    void set_size(int n, bool = false) { d.resize(n); }
    void set_length(int n, bool = false) { d.resize(n); }
    void clear() { for (size_t i = 0; i < d.size(); ++i) d[i] = bin(0); }
    // IT++ release builds do not bounds-check; the reference reads final[36..47] of a 36-element vector
    // (lib/recc_decode_impl.cc:68,71-77) and ignores what it gets, so out-of-range reads land on a scratch element.
    bin &operator[](int i) { return (i >= 0 && i < (int)d.size()) ? d[i] : scratch(); }
    const bin &operator[](int i) const { return (i >= 0 && i < (int)d.size()) ? d[i] : scratch(); }
    bin &operator()(int i) { return (*this)[i]; }
    const bin &operator()(int i) const { return (*this)[i]; }
    bvec operator()(int i1, int i2) const {          // elements i1..i2 inclusive
        bvec r;
        for (int i = i1; i <= i2 && i < (int)d.size(); ++i) r.d.push_back(d[i]);
        return r;
    }
    bvec mid(int start, int n) const { return (*this)(start, start + n - 1); }
    void replace_mid(int pos, const bvec &v) { for (int i = 0; i < v.size(); ++i) (*this)[pos + i] = v[i]; }
    void push(bin b) { d.push_back(b); }
private:
    static bin &scratch() { static bin s; s = bin(0); return s; }
    void parse(const char *s) {
        std::istringstream is(s);
        int v;
        while (is >> v) d.push_back(bin(v));
    }
    std::vector<bin> d;
};

inline bvec concat(const bvec &a, const bvec &b) {
    bvec r(a);
    for (int i = 0; i < b.size(); ++i) r.push(b[i]);
    return r;
}
inline std::ostream &operator<<(std::ostream &os, const bvec &v) {
    os << "[";
    for (int i = 0; i < v.size(); ++i) os << (i ? " " : "") << (int)v[i];
    return os << "]";
}

class BCH {
public:
    BCH(int in_n, int in_t, bool sys = false) : n(in_n), t(in_t), systematic(sys) {
        m = 0;
        while ((1 << m) - 1 < n) ++m;
        if ((1 << m) - 1 != n || m < 3 || m > 8) { std::cerr << "itpp stand-in: BCH length must be 2^m-1, m=3..8\n"; abort(); }
        static const int prim[9] = {0, 0, 0, 0xB, 0x13, 0x25, 0x43, 0x89, 0x11D};
        alog.assign(n, 0);
        logt.assign(n + 1, -1);
        int x = 1;
        for (int i = 0; i < n; ++i) {
            alog[i] = x;
            logt[x] = i;
            x <<= 1;
            if (x & (1 << m)) x ^= prim[m];
        }
        // generator: product of the distinct minimal polynomials of alpha^1..alpha^(2t) (coefficients end up in GF(2))
        std::vector<int> g1(1, 0);               // the polynomial "1": exponent notation, g1[i] = coefficient of x^i
        std::vector<char> used(n, 0);
        for (int j = 1; j <= 2 * t; ++j) {
            if (used[j % n]) continue;
            int e = j % n;
            do {                                 // multiply by (x + alpha^e) over the whole cyclotomic coset of j
                used[e] = 1;
                std::vector<int> nx(g1.size() + 1, -1);
                for (size_t i = 0; i < g1.size(); ++i) {
                    nx[i + 1] = add(nx[i + 1], g1[i]);
                    nx[i] = add(nx[i], mul(g1[i], e));
                }
                g1.swap(nx);
                e = (2 * e) % n;
            } while (e != j % n);
        }
        g = g1;
        for (size_t i = 0; i < g.size(); ++i)
            if (g[i] > 0) { std::cerr << "itpp stand-in: generator not binary\n"; abort(); }
        k = n - ((int)g.size() - 1);
    }
    int get_k() const { return k; }

    bvec encode(const bvec &uncoded) {
        bvec coded;
        encode(uncoded, coded);
        return coded;
    }
    void encode(const bvec &uncoded, bvec &coded) {
        const int iterations = uncoded.length() / k;
        coded.set_size(iterations * n);
        for (int it = 0; it < iterations; ++it) {
            bvec mbit = uncoded.mid(it * k, k);
            std::vector<int> c(n, -1);           // c[j] = coefficient of x^j
            if (systematic) {
                for (int j = 0; j < k; ++j) c[j + n - k] = (int)mbit(k - j - 1) - 1;
                std::vector<int> r = mod_g(c);
                for (int j = 0; j < n - k; ++j) c[j] = r[j];
            } else {
                std::vector<int> mp(k, -1);
                for (int j = 0; j < k; ++j) mp[j] = (int)mbit(k - j - 1) - 1;
                for (int i = 0; i < k; ++i)
                    for (size_t j = 0; j < g.size(); ++j) c[i + j] = add(c[i + j], mul(mp[i], g[j]));
            }
            for (int j = 0; j < n; ++j) coded(it * n + j) = bin(c[n - j - 1] == 0 ? 1 : 0);
        }
    }

    bool decode(const bvec &coded, bvec &decoded, bvec &cw_isvalid) {
        const int iterations = coded.length() / n;
        decoded.set_size(iterations * k);
        cw_isvalid.set_length(iterations);
        bool no_dec_failure = true;
        for (int it = 0; it < iterations; ++it) {
            bool failure = false;
            bvec rbin = coded.mid(it * n, n);
            std::vector<int> r(n), c(n);
            for (int j = 0; j < n; ++j) r[j] = (int)rbin(n - j - 1) - 1;
            std::vector<int> S(2 * t + 1, -1);
            for (int j = 1; j <= 2 * t; ++j) S[j] = eval(r, j);
            if (true_degree(S) >= 1) {
                std::vector<int> Sp1(S);
                Sp1[0] = 0;                       // S(x) + 1
                std::vector<int> Lambda(1, 0), T(1, 0);
                for (int kk = 0; kk < t; ++kk) {
                    std::vector<int> Omega = pmul(Lambda, Sp1);
                    const int delta = (2 * kk + 1 < (int)Omega.size()) ? Omega[2 * kk + 1] : -1;
                    std::vector<int> Old(Lambda);
                    Lambda = padd(Old, pscale(pshift(T, 1), delta));
                    if (delta == -1 || true_degree(Old) > kk) T = pshift(T, 2);
                    else T = pscale(pshift(Old, 1), delta < 0 ? -1 : (n - delta) % n);
                }
                const int deg = true_degree(Lambda);
                std::vector<int> errorpos;
                for (int j = 0; j <= n - 1 && deg > 0; ++j) {
                    if (eval(Lambda, j) == -1) {
                        errorpos.push_back((n - j) % n);
                        if ((int)errorpos.size() >= deg) break;
                    }
                }
                if ((int)errorpos.size() != deg) {
                    failure = true;
                } else {
                    for (size_t j = 0; j < errorpos.size(); ++j) rbin(n - errorpos[j] - 1) += bin(1);
                    for (int j = 0; j < n; ++j) c[j] = (int)rbin(n - j - 1) - 1;
                    std::vector<int> S2(2 * t + 1, -1);
                    for (int j = 1; j <= 2 * t; ++j) S2[j] = eval(c, j);
                    failure = true_degree(S2) > 0;
                }
            } else {
                c = r;
            }
            bvec mbit(k);
            if (!failure) {
                if (systematic) {
                    for (int j = 0; j < k; ++j) mbit(k - j - 1) = bin(c[n - k + j] == -1 ? 0 : 1);
                } else {
                    std::vector<int> q = div_g(c);
                    for (int j = 0; j < k; ++j) mbit(k - j - 1) = bin(j < (int)q.size() && q[j] != -1 ? 1 : 0);
                }
            } else {
                if (systematic) mbit = coded.mid(it * n, k);
                no_dec_failure = false;
            }
            decoded.replace_mid(it * k, mbit);
            cw_isvalid(it) = bin(failure ? 0 : 1);
        }
        return no_dec_failure;
    }

private:
    // field elements are exponents of alpha; -1 is the zero element (IT++'s GF convention)
    int mul(int a, int b) const { return (a < 0 || b < 0) ? -1 : (a + b) % n; }
    int add(int a, int b) const {
        const int v = (a < 0 ? 0 : alog[a]) ^ (b < 0 ? 0 : alog[b]);
        return v ? logt[v] : -1;
    }
    static int true_degree(const std::vector<int> &p) {
        for (int i = (int)p.size() - 1; i >= 0; --i) if (p[i] != -1) return i;
        return -1;
    }
    int eval(const std::vector<int> &p, int e) const {      // p(alpha^e), Horner
        int acc = -1;
        for (int i = (int)p.size() - 1; i >= 0; --i) acc = add(mul(acc, e % n), p[i]);
        return acc;
    }
    std::vector<int> pmul(const std::vector<int> &a, const std::vector<int> &b) const {
        std::vector<int> r(a.size() + b.size() - 1, -1);
        for (size_t i = 0; i < a.size(); ++i)
            for (size_t j = 0; j < b.size(); ++j) r[i + j] = add(r[i + j], mul(a[i], b[j]));
        return r;
    }
    std::vector<int> padd(const std::vector<int> &a, const std::vector<int> &b) const {
        std::vector<int> r(a.size() > b.size() ? a.size() : b.size(), -1);
        for (size_t i = 0; i < r.size(); ++i) r[i] = add(i < a.size() ? a[i] : -1, i < b.size() ? b[i] : -1);
        return r;
    }
    std::vector<int> pscale(const std::vector<int> &a, int e) const {
        std::vector<int> r(a);
        for (size_t i = 0; i < r.size(); ++i) r[i] = mul(r[i], e);
        return r;
    }
    static std::vector<int> pshift(const std::vector<int> &a, int s) {
        std::vector<int> r(a.size() + s, -1);
        for (size_t i = 0; i < a.size(); ++i) r[i + s] = a[i];
        return r;
    }
    std::vector<int> mod_g(std::vector<int> a) const {      // a(x) mod g(x); g is monic with binary coefficients
        const int dg = (int)g.size() - 1;
        for (int i = (int)a.size() - 1; i >= dg; --i) {
            const int q = a[i];
            if (q == -1) continue;
            for (int j = 0; j <= dg; ++j) a[i - dg + j] = add(a[i - dg + j], mul(q, g[j]));
        }
        a.resize(dg);
        return a;
    }
    std::vector<int> div_g(std::vector<int> a) const {
        const int dg = (int)g.size() - 1;
        std::vector<int> q(a.size() > (size_t)dg ? a.size() - dg : 1, -1);
        for (int i = (int)a.size() - 1; i >= dg; --i) {
            const int c = a[i];
            if (c == -1) continue;
            q[i - dg] = c;
            for (int j = 0; j <= dg; ++j) a[i - dg + j] = add(a[i - dg + j], mul(c, g[j]));
        }
        return q;
    }

    int n, t, k, m;
    bool systematic;
    std::vector<int> alog, logt, g;
};

}  // namespace itpp
