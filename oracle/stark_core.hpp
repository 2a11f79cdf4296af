/* ORACLE (test infrastructure, NOT product code) -- see oracle/gl.h header.
 *
 * CPU restatement of the transcript, the FRI prover/verifier and the single-table / multi-table STARK
 * prover and verifier of the reference.  Deliberately simple (scalar, P::WIDTH = 1 semantics).
 *
 * Follows:
 *   plonky2/plonky2/src/iop/challenger.rs:18-162            Challenger (duplex sponge, pops from the END :97-99)
 *   plonky2/plonky2/src/fri/challenges.rs:15-76             observe_openings, fri_challenges
 *   plonky2/plonky2/src/fri/oracle.rs:167-241               prove_openings
 *   plonky2/plonky2/src/fri/prover.rs:20-204                fri_proof, fri_committed_trees, PoW, queries
 *   plonky2/plonky2/src/fri/verifier.rs:18-262              verify_fri_proof
 *   plonky2/plonky2/src/fri/reduction_strategies.rs:40-53   ConstantArityBits(4, 5)
 *   plonky2/plonky2/src/util/reducing.rs:27-100             ReducingFactor
 *   plonky2/field/src/polynomial/division.rs:74-87          divide_by_linear
 *   circuits/src/stark/prover.rs:79-705                     prove_with_traces, prove_single_table, compute_quotient_polys
 *   circuits/src/stark/{constraint_consumer,vanishing_poly,cross_table_lookup,permutation,proof,
 *                       get_challenges,verifier,serialization,config,stark}.rs
 */
#ifndef ORC_STARK_CORE_HPP
#define ORC_STARK_CORE_HPP
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "oracle.h"

namespace orc {

typedef uint64_t F;
typedef gl2_t E;
typedef std::vector<F> VF;
typedef std::vector<E> VE;

struct Hash {
    F e[4];
    bool operator==(const Hash& o) const { return memcmp(e, o.e, 32) == 0; }
};
typedef std::vector<Hash> Cap;

inline E e_from(F a) { return gl2_make(a, 0); }
inline E e_zero() { return gl2_make(0, 0); }
inline E e_one() { return gl2_make(1, 0); }

/* field-generic helpers so that AIR constraints are written once over P in {F, E} */
struct FOps {
    typedef F T;
    static T add(T a, T b) { return gl_add(a, b); }
    static T sub(T a, T b) { return gl_sub(a, b); }
    static T mul(T a, T b) { return gl_mul(a, b); }
    static T from(F a) { return a; }
};
struct EOps {
    typedef E T;
    static T add(T a, T b) { return gl2_add(a, b); }
    static T sub(T a, T b) { return gl2_sub(a, b); }
    static T mul(T a, T b) { return gl2_mul(a, b); }
    static T from(F a) { return e_from(a); }
};
/* thin value wrapper with operators, P<FOps> / P<EOps> */
template <class O>
struct P {
    typename O::T v;
    P() : v(O::from(0)) {}
    P(typename O::T x) : v(x) {}
    static P c(uint64_t k) { return P(O::from(k % GL_P)); }
    static P one() { return c(1); }
    static P zero() { return c(0); }
    P operator+(P o) const { return P(O::add(v, o.v)); }
    P operator-(P o) const { return P(O::sub(v, o.v)); }
    P operator*(P o) const { return P(O::mul(v, o.v)); }
    P operator*(F s) const { return P(O::mul(v, O::from(s))); }
    P& operator+=(P o) { v = O::add(v, o.v); return *this; }
    P& operator-=(P o) { v = O::sub(v, o.v); return *this; }
    P& operator*=(P o) { v = O::mul(v, o.v); return *this; }
};

/* ------------------------------------------------------------------ Challenger (challenger.rs) */
struct Challenger {
    F state[12];
    VF in, out;
    Challenger() { memset(state, 0, sizeof(state)); }
    void duplexing() { /* :137-152 */
        for (size_t i = 0; i < in.size(); i++) state[i] = in[i];
        in.clear();
        orc_challenger_permute(state); /* H::Permutation: Poseidon, or Blake3Permutation (hash/blake3.rs:165-199) */
        out.assign(state, state + 8);
    }
    void observe(F x) { /* :47-56 */
        out.clear();
        in.push_back(gl_canon(x));
        if (in.size() == 8) duplexing();
    }
    void observe_ext(E x) { observe(x.c0); observe(x.c1); }
    void observe_hash(const Hash& h) { /* :79-81 observes hash.to_vec(): HashOut -> its 4 elements; BytesHash<32> -> 5
                                          elements of 7 bytes each (hash_types.rs:142-152) */
        if (orc_get_hasher() == 1) {
            F f[5];
            orc_bytes_hash_to_fields(h.e, f);
            for (int i = 0; i < 5; i++) observe(f[i]);
        } else
            for (int i = 0; i < 4; i++) observe(h.e[i]);
    }
    void observe_cap(const Cap& c) { for (auto& h : c) observe_hash(h); }
    F get_challenge() { /* :86-99 */
        if (!in.empty() || out.empty()) duplexing();
        F r = out.back();
        out.pop_back();
        return r;
    }
    VF get_n(size_t n) { VF r; for (size_t i = 0; i < n; i++) r.push_back(get_challenge()); return r; }
    E get_ext() { F a = get_challenge(); F b = get_challenge(); return gl2_make(a, b); }
    Hash get_hash() { Hash h; for (int i = 0; i < 4; i++) h.e[i] = get_challenge(); return h; }
    void compact() { /* :154-160 */
        if (!in.empty()) duplexing();
        out.clear();
    }
};

/* ------------------------------------------------------------------ config (config.rs:18-30) */
struct Config {
    uint32_t num_challenges = 2, rate_bits = 3, cap_height = 4, pow_bits = 16, arity_bits = 4, final_poly_bits = 5,
             num_queries = 28;
    bool check_quotient_degree = true; /* false: "pipeline parity" on non-satisfying traces */
};
struct FriParams {
    uint32_t degree_bits;
    std::vector<uint32_t> arity_bits;
    uint32_t total_arities() const { uint32_t s = 0; for (auto a : arity_bits) s += a; return s; }
};
inline FriParams fri_params(const Config& c, uint32_t degree_bits) { /* reduction_strategies.rs:40-53 */
    FriParams p;
    p.degree_bits = degree_bits;
    uint32_t d = degree_bits;
    while (d > c.final_poly_bits && d + c.rate_bits - c.arity_bits >= c.cap_height) {
        p.arity_bits.push_back(c.arity_bits);
        d -= c.arity_bits;
    }
    return p;
}

/* ------------------------------------------------------------------ PolynomialBatch (fri/oracle.rs) */
struct Batch {
    size_t ncols = 0, n = 0, L = 0;
    uint32_t cap_height = 0;
    VF coeffs;  /* [ncols][n] */
    VF leaves;  /* [L][ncols] */
    VF digests; /* reference layout */
    Cap cap;
    const F* leaf(size_t i) const { return &leaves[i * ncols]; }
    std::vector<Hash> prove(size_t i) const {
        uint32_t nl = orc_log2_strict(L) - cap_height;
        std::vector<Hash> s(nl);
        if (nl) orc_merkle_prove(digests.data(), L, cap_height, i, (uint64_t*)s.data());
        return s;
    }
};
inline Batch commit(const VF& cols, size_t ncols, size_t n, bool is_coeffs, const Config& c) {
    Batch b;
    b.ncols = ncols;
    b.n = n;
    b.L = n << c.rate_bits;
    b.cap_height = c.cap_height;
    b.coeffs.resize(ncols * n);
    b.leaves.resize(b.L * ncols);
    size_t ncap = (size_t)1 << c.cap_height;
    b.digests.resize(4 * (2 * (b.L - ncap) + 1));
    b.cap.resize(ncap);
    int rc = orc_commit(cols.data(), ncols, n, is_coeffs, c.rate_bits, c.cap_height, b.coeffs.data(), b.leaves.data(),
                        b.digests.data(), (uint64_t*)b.cap.data());
    if (rc) throw std::runtime_error("commit failed");
    return b;
}

/* tree over arbitrary row-major leaves (FRI layers) */
struct Tree {
    size_t nrows = 0, ncols = 0;
    uint32_t cap_height = 0;
    VF leaves, digests;
    Cap cap;
    std::vector<Hash> prove(size_t i) const {
        uint32_t nl = orc_log2_strict(nrows) - cap_height;
        std::vector<Hash> s(nl);
        if (nl) orc_merkle_prove(digests.data(), nrows, cap_height, i, (uint64_t*)s.data());
        return s;
    }
};
inline Tree build_tree(VF leaves, size_t nrows, size_t ncols, uint32_t cap_height) {
    Tree t;
    t.nrows = nrows;
    t.ncols = ncols;
    t.cap_height = cap_height;
    size_t ncap = (size_t)1 << cap_height;
    t.digests.resize(4 * (2 * (nrows - ncap) + 1));
    t.cap.resize(ncap);
    if (orc_merkle_new_v2(leaves.data(), nrows, ncols, cap_height, t.digests.data(), (uint64_t*)t.cap.data()))
        throw std::runtime_error("merkle failed");
    t.leaves = std::move(leaves);
    return t;
}

/* ------------------------------------------------------------------ proof structures (proof.rs, fri/proof.rs) */
struct FriQueryStep { VE evals; std::vector<Hash> siblings; };
struct FriQueryRound {
    std::vector<std::pair<VF, std::vector<Hash>>> initial; /* per oracle: (leaf, merkle proof) */
    std::vector<FriQueryStep> steps;
};
struct FriProof {
    std::vector<Cap> commit_caps;
    std::vector<FriQueryRound> rounds;
    VE final_poly;
    F pow_witness = 0;
};
struct OpeningSet { /* proof.rs:181-196 */
    VE local_values, next_values, zs, zs_next;
    VF ctl_zs_last;
    VE quotient;
};
struct StarkProof {
    Cap trace_cap, zs_cap, quotient_cap;
    OpeningSet openings;
    FriProof fri;
};

/* FRI instance: batches of (point, list of (oracle, poly)) -- stark.rs:87-150 */
struct FriBatch { E point; std::vector<std::pair<int, int>> polys; };
struct FriInstance { std::vector<size_t> oracle_num_polys; std::vector<FriBatch> batches; };

/* ------------------------------------------------------------------ ext polynomial helpers */
inline E poly_eval_ext_base(const F* c, size_t n, E x) { /* PolynomialCoeffs::to_extension().eval */
    E acc = e_zero();
    for (size_t i = n; i-- > 0;) acc = gl2_add(gl2_mul(acc, x), e_from(c[i]));
    return acc;
}
inline E poly_eval_ext(const VE& c, E x) {
    E acc = e_zero();
    for (size_t i = c.size(); i-- > 0;) acc = gl2_add(gl2_mul(acc, x), c[i]);
    return acc;
}
/* coset FFT of an extension polynomial = the base transform on each component (the 2^k-th roots of unity
 * and the shift live in the base field: goldilocks_extensions.rs:27, types.rs:240) */
inline VE ext_coset_fft(const VE& coeffs, F shift) {
    size_t n = coeffs.size();
    VF a(n), b(n), oa(n), ob(n);
    for (size_t i = 0; i < n; i++) { a[i] = coeffs[i].c0; b[i] = coeffs[i].c1; }
    orc_evaluate_poly_with_offset(a.data(), n, shift, 1, oa.data());
    orc_evaluate_poly_with_offset(b.data(), n, shift, 1, ob.data());
    VE r(n);
    for (size_t i = 0; i < n; i++) r[i] = gl2_make(oa[i], ob[i]);
    return r;
}
template <class T>
inline void reverse_index_bits(std::vector<T>& v) {
    size_t n = v.size();
    uint32_t lg = orc_log2_strict(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = orc_bitrev(i, lg);
        if (j > i) std::swap(v[i], v[j]);
    }
}

/* ------------------------------------------------------------------ FRI prover (oracle.rs:167-241, prover.rs) */
inline F fri_pow(const Hash& h, const Config& c) { /* smallest valid nonce (the reference's
                                                      rayon find_any returns *some* valid nonce) */
    for (uint64_t i = 0;; i++) {
        F in[5] = {h.e[0], h.e[1], h.e[2], h.e[3], i};
        F out[4];
        orc_poseidon_hash_no_pad(in, 5, out); /* C::InnerHasher = PoseidonHash in every config (plonk/config.rs:121,159) */
        if (__builtin_clzll(out[0] | 1) >= (int)c.pow_bits && (out[0] >> (64 - c.pow_bits)) == 0) return i;
    }
}
inline bool pow_ok(F response, const Config& c) { return (response >> (64 - c.pow_bits)) == 0; }

inline FriProof prove_openings(const FriInstance& inst, const std::vector<const Batch*>& oracles, Challenger& ch,
                               const FriParams& fp, const Config& cfg) {
    E alpha = ch.get_ext();
    size_t n = oracles[0]->n;
    VE final_poly; /* empty */
    for (auto& b : inst.batches) {
        size_t len = b.polys.size();
        VE comp(n, e_zero());
        E ap = e_one();
        for (size_t i = 0; i < len; i++) {
            const Batch* o = oracles[b.polys[i].first];
            const F* c = &o->coeffs[(size_t)b.polys[i].second * n];
            for (size_t j = 0; j < n; j++) comp[j] = gl2_add(comp[j], gl2_scalar_mul(ap, c[j]));
            ap = gl2_mul(ap, alpha);
        }
        /* divide_by_linear (division.rs:74-87) */
        VE q(n);
        E acc = e_zero();
        for (size_t k = n; k-- > 0;) {
            acc = gl2_add(gl2_mul(acc, b.point), comp[k]);
            q[k] = acc;
        }
        VE quot(q.begin() + 1, q.end()); /* bs.pop(); bs.reverse() */
        E shift = gl2_pow(alpha, len);
        for (auto& x : final_poly) x = gl2_mul(x, shift);
        if (final_poly.size() < quot.size()) final_poly.resize(quot.size(), e_zero());
        for (size_t j = 0; j < quot.size(); j++) final_poly[j] = gl2_add(final_poly[j], quot[j]);
    }
    final_poly.insert(final_poly.begin(), e_zero());
    VE coeffs = final_poly;
    coeffs.resize(n << cfg.rate_bits, e_zero());
    VE values = ext_coset_fft(coeffs, GL_GEN);

    FriProof pr;
    std::vector<Tree> trees;
    F shift = GL_GEN;
    for (uint32_t ab : fp.arity_bits) { /* fri_committed_trees */
        size_t arity = (size_t)1 << ab;
        reverse_index_bits(values);
        size_t nl = values.size() / arity;
        VF leaves(values.size() * 2);
        for (size_t i = 0; i < values.size(); i++) { leaves[2 * i] = values[i].c0; leaves[2 * i + 1] = values[i].c1; }
        Tree t = build_tree(std::move(leaves), nl, arity * 2, cfg.cap_height);
        ch.observe_cap(t.cap);
        pr.commit_caps.push_back(t.cap);
        trees.push_back(std::move(t));
        E beta = ch.get_ext();
        VE nc(coeffs.size() / arity);
        for (size_t j = 0; j < nc.size(); j++) { /* reduce_with_powers */
            E s = e_zero();
            for (size_t i = arity; i-- > 0;) s = gl2_add(gl2_mul(s, beta), coeffs[j * arity + i]);
            nc[j] = s;
        }
        coeffs = nc;
        shift = gl_pow(shift, arity);
        values = ext_coset_fft(coeffs, shift);
    }
    coeffs.resize(coeffs.size() >> cfg.rate_bits);
    for (auto& c : coeffs) ch.observe_ext(c);
    pr.final_poly = coeffs;
    Hash h = ch.get_hash();
    pr.pow_witness = fri_pow(h, cfg);
    size_t Lsz = n << cfg.rate_bits;
    VF qs = ch.get_n(cfg.num_queries);
    for (F r : qs) {
        size_t x = (size_t)(r % Lsz);
        FriQueryRound qr;
        for (auto* o : oracles) qr.initial.push_back({VF(o->leaf(x), o->leaf(x) + o->ncols), o->prove(x)});
        for (size_t i = 0; i < trees.size(); i++) {
            uint32_t ab = fp.arity_bits[i];
            size_t idx = x >> ab;
            FriQueryStep st;
            const F* lf = &trees[i].leaves[idx * trees[i].ncols];
            for (size_t k = 0; k < trees[i].ncols / 2; k++) st.evals.push_back(gl2_make(lf[2 * k], lf[2 * k + 1]));
            st.siblings = trees[i].prove(idx);
            qr.steps.push_back(std::move(st));
            x >>= ab;
        }
        pr.rounds.push_back(std::move(qr));
    }
    return pr;
}

/* ------------------------------------------------------------------ FRI verifier (fri/verifier.rs) */
struct FriChallenges { E alpha; VE betas; F pow_response; std::vector<size_t> indices; };
inline FriChallenges fri_challenges(Challenger& ch, const FriProof& p, uint32_t degree_bits, const Config& c) { /* challenges.rs:25-75 */
    FriChallenges fc;
    fc.alpha = ch.get_ext();
    for (auto& cap : p.commit_caps) { ch.observe_cap(cap); fc.betas.push_back(ch.get_ext()); }
    for (auto& x : p.final_poly) ch.observe_ext(x);
    Hash h = ch.get_hash();
    F in[5] = {h.e[0], h.e[1], h.e[2], h.e[3], p.pow_witness};
    F out[4];
    orc_poseidon_hash_no_pad(in, 5, out); /* C::InnerHasher = PoseidonHash in every config (plonk/config.rs:121,159) */
    fc.pow_response = out[0];
    size_t L = (size_t)1 << (degree_bits + c.rate_bits);
    for (uint32_t i = 0; i < c.num_queries; i++) fc.indices.push_back((size_t)(ch.get_challenge() % L));
    return fc;
}
inline E interpolate_eval(const VE& xs, const VE& ys, E x) { /* unique interpolant (interpolation.rs) */
    E r = e_zero();
    for (size_t i = 0; i < xs.size(); i++) {
        E num = ys[i], den = e_one();
        for (size_t j = 0; j < xs.size(); j++)
            if (j != i) { num = gl2_mul(num, gl2_sub(x, xs[j])); den = gl2_mul(den, gl2_sub(xs[i], xs[j])); }
        r = gl2_add(r, gl2_mul(num, gl2_inv(den)));
    }
    return r;
}
inline std::string verify_fri(const FriInstance& inst, const std::vector<VE>& openings, const FriChallenges& fc,
                              const std::vector<Cap>& initial_caps, const FriProof& p, const FriParams& fp, const Config& c) {
    size_t lde_bits = fp.degree_bits + c.rate_bits, n = (size_t)1 << lde_bits;
    if (!pow_ok(fc.pow_response, c)) return "Invalid proof of work witness.";
    if (p.rounds.size() != c.num_queries) return "Number of query rounds does not match config.";
    if (p.final_poly.size() != ((size_t)1 << (fp.degree_bits - fp.total_arities()))) return "final poly shape";
    VE reduced; /* PrecomputedReducedOpenings */
    for (auto& b : openings) {
        E s = e_zero();
        for (size_t i = b.size(); i-- > 0;) s = gl2_add(gl2_mul(s, fc.alpha), b[i]);
        reduced.push_back(s);
    }
    for (size_t qi = 0; qi < p.rounds.size(); qi++) {
        size_t x = fc.indices[qi];
        const FriQueryRound& qr = p.rounds[qi];
        if (qr.initial.size() != initial_caps.size()) return "initial shape";
        for (size_t o = 0; o < qr.initial.size(); o++) {
            if (qr.initial[o].first.size() != inst.oracle_num_polys[o]) return "leaf shape";
            if (qr.initial[o].second.size() + c.cap_height != lde_bits) return "path shape";
            if (!orc_merkle_verify(qr.initial[o].first.data(), qr.initial[o].first.size(), x, (const uint64_t*)initial_caps[o].data(),
                                   (const uint64_t*)qr.initial[o].second.data(), qr.initial[o].second.size()))
                return "Invalid Merkle proof (initial).";
        }
        F sx = gl_mul(GL_GEN, gl_pow(gl_root_of_unity((int)lde_bits), orc_bitrev(x, lde_bits)));
        /* fri_combine_initial :117-160 */
        E sum = e_zero();
        E subgroup_x = e_from(sx);
        for (size_t bi = 0; bi < inst.batches.size(); bi++) {
            auto& b = inst.batches[bi];
            E red = e_zero();
            for (size_t i = b.polys.size(); i-- > 0;)
                red = gl2_add(gl2_mul(red, fc.alpha), e_from(qr.initial[b.polys[i].first].first[b.polys[i].second]));
            E num = gl2_sub(red, reduced[bi]);
            E den = gl2_sub(subgroup_x, b.point);
            sum = gl2_mul(sum, gl2_pow(fc.alpha, b.polys.size())); /* alpha.shift: count == #polys just reduced */
            sum = gl2_add(sum, gl2_mul(num, gl2_inv(den)));
        }
        E old_eval = gl2_mul(sum, subgroup_x);
        if (qr.steps.size() != fp.arity_bits.size()) return "steps shape";
        size_t bits = lde_bits;
        for (size_t i = 0; i < fp.arity_bits.size(); i++) {
            uint32_t ab = fp.arity_bits[i];
            size_t arity = (size_t)1 << ab;
            const VE& ev = qr.steps[i].evals;
            if (ev.size() != arity) return "evals shape";
            size_t coset_index = x >> ab, within = x & (arity - 1);
            if (!gl2_eq(ev[within], old_eval)) return "FRI consistency check failed.";
            /* compute_evaluation :18-41 */
            F g = gl_root_of_unity((int)ab);
            VE evs = ev;
            reverse_index_bits(evs);
            size_t rev = orc_bitrev(within, ab);
            F start = gl_mul(sx, gl_pow(g, arity - rev));
            VE xs;
            F y = 1;
            for (size_t k = 0; k < arity; k++) { xs.push_back(e_from(gl_mul(start, y))); y = gl_mul(y, g); }
            old_eval = interpolate_eval(xs, evs, fc.betas[i]);
            VF flat;
            for (auto& e : ev) { flat.push_back(e.c0); flat.push_back(e.c1); }
            bits -= ab;
            if (qr.steps[i].siblings.size() + c.cap_height != bits) return "step path shape";
            if (!orc_merkle_verify(flat.data(), flat.size(), coset_index, (const uint64_t*)p.commit_caps[i].data(),
                                   (const uint64_t*)qr.steps[i].siblings.data(), qr.steps[i].siblings.size()))
                return "Invalid Merkle proof (step).";
            for (uint32_t k = 0; k < ab; k++) sx = gl_sqr(sx);
            x = coset_index;
        }
        if (!gl2_eq(poly_eval_ext(p.final_poly, e_from(sx)), old_eval)) return "Final polynomial evaluation is invalid.";
    }
    return "";
}

}  // namespace orc
#endif
