/* ORACLE (test infrastructure, NOT product code) -- see oracle/stark_core.hpp for the reference map.
 * Multi-table STARK prover / verifier / wire format, generic over the AIR tables (tables.hpp). */
#ifndef ORC_STARK_HPP
#define ORC_STARK_HPP
#include "stark_core.hpp"

namespace orc {

/* ------------------------------------------------------------------ ConstraintConsumer (constraint_consumer.rs:10-80) */
template <class O>
struct Consumer {
    typedef P<O> T;
    std::vector<F> alphas;
    std::vector<T> accs;
    T z_last, lagrange_first, lagrange_last;
    Consumer(const VF& a, T zl, T lf, T ll) : alphas(a), accs(a.size(), T::zero()), z_last(zl), lagrange_first(lf), lagrange_last(ll) {}
    /* test hook (orc_air_first_failure): position, in evaluation order, of the first constraint that is non-zero */
    int seen = 0, first_nonzero = -1;
    static bool is_zero(F x) { return gl_canon(x) == 0; }
    static bool is_zero(E x) { return gl_canon(x.c0) == 0 && gl_canon(x.c1) == 0; }
    /* test hook (orc_air_constraints): the raw argument and the kind of every yield, in order */
    bool record = false;
    std::vector<typename O::T> rec_raw;
    std::vector<int> rec_kinds;
    void note(T c, int kind) {
        if (record) {
            rec_raw.push_back(c.v);
            rec_kinds.push_back(kind);
        }
    }
    void weigh(T c) {
        if (first_nonzero < 0 && !is_zero(c.v)) first_nonzero = seen;
        seen++;
        for (size_t i = 0; i < alphas.size(); i++) accs[i] = accs[i] * alphas[i] + c;
    }
    void constraint(T c) { note(c, 0); weigh(c); }
    void constraint_transition(T c) { note(c, 1); weigh(c * z_last); }
    void constraint_first_row(T c) { note(c, 2); weigh(c * lagrange_first); }
    void constraint_last_row(T c) { note(c, 3); weigh(c * lagrange_last); }
};

/* ------------------------------------------------------------------ Column / CTL (cross_table_lookup.rs) */
struct Column {
    std::vector<std::pair<int, F>> lc;
    F constant = 0;
    static Column single(int c) { Column r; r.lc.push_back({c, 1}); return r; }
    static Column constant_(F k) { Column r; r.constant = k; return r; }
    static Column linear(std::vector<std::pair<int, F>> v, F k = 0) { Column r; r.lc = std::move(v); r.constant = k; return r; }
    static Column sum(const std::vector<int>& cs) { Column r; for (int c : cs) r.lc.push_back({c, 1}); return r; }
    static Column le_bits(const std::vector<int>& cs) { Column r; F w = 1; for (int c : cs) { r.lc.push_back({c, w}); w = gl_add(w, w); } return r; }
    template <class O>
    P<O> eval(const P<O>* v) const { /* :102-110 */
        P<O> s = P<O>::zero();
        for (auto& t : lc) s = s + v[t.first] * t.second;
        return s + P<O>::c(constant);
    }
    F eval_table(const VF& trace, size_t n, size_t row) const { /* :112-118 */
        F s = 0;
        for (auto& t : lc) s = gl_add(s, gl_mul(trace[(size_t)t.first * n + row], t.second));
        return gl_add(s, constant);
    }
};
inline std::vector<Column> singles(const std::vector<int>& cs) { std::vector<Column> r; for (int c : cs) r.push_back(Column::single(c)); return r; }

struct TableWithColumns { int table; std::vector<Column> columns; bool has_filter = false; Column filter; };
inline TableWithColumns twc(int table, std::vector<Column> cols, Column filter) { TableWithColumns t; t.table = table; t.columns = std::move(cols); t.has_filter = true; t.filter = std::move(filter); return t; }
inline TableWithColumns twc_nofilter(int table, std::vector<Column> cols) { TableWithColumns t; t.table = table; t.columns = std::move(cols); return t; }
struct CrossTableLookup { std::vector<TableWithColumns> looking; TableWithColumns looked; bool has_looked = true; bool missing_sides = false; /* a side's table has no restatement yet */ bool complete = true; /* every side of the registered CTL is inside the system */ };

struct Challenge { F beta, gamma; };
struct CtlZ { VF z; Challenge ch; std::vector<Column> columns; bool has_filter; Column filter; };

template <class O>
P<O> combine(const Challenge& c, const std::vector<P<O>>& terms) { /* permutation.rs:61-72: reduce_with_powers + gamma */
    P<O> s = P<O>::zero();
    for (size_t i = terms.size(); i-- > 0;) s = s * c.beta + terms[i];
    return s + P<O>::c(c.gamma);
}

/* partial_products (cross_table_lookup.rs:284-310).  Returns "" or an error (non-binary filter). */
inline std::string partial_products(const VF& trace, size_t n, const TableWithColumns& t, Challenge ch, VF& out) {
    out.resize(n);
    F prod = 1;
    for (size_t i = 0; i < n; i++) {
        F filter = t.has_filter ? t.filter.eval_table(trace, n, i) : 1;
        if (filter == 1) {
            std::vector<P<FOps>> ev;
            for (auto& c : t.columns) ev.push_back(P<FOps>(c.eval_table(trace, n, i)));
            prod = gl_mul(prod, combine<FOps>(ch, ev).v);
        } else if (filter != 0) {
            return "Non-binary filter?";
        }
        out[i] = prod;
    }
    return "";
}

/* ------------------------------------------------------------------ Stark table interface (stark.rs) */
struct PermutationPair { std::vector<std::pair<int, int>> column_pairs; };
struct Table {
    std::string name;
    int columns = 0;
    int constraint_degree = 0;
    std::vector<PermutationPair> permutation_pairs;
    std::function<void(const P<FOps>*, const P<FOps>*, Consumer<FOps>&)> eval_base;
    std::function<void(const P<EOps>*, const P<EOps>*, Consumer<EOps>&)> eval_ext;
    int quotient_degree_factor() const { return std::max(1, constraint_degree - 1); }
    bool uses_permutation_args() const { return !permutation_pairs.empty(); }
    int permutation_batch_size() const { return quotient_degree_factor(); }
    int num_permutation_batches(const Config& c) const {
        int inst = (int)permutation_pairs.size() * (int)c.num_challenges;
        int bs = permutation_batch_size();
        return (inst + bs - 1) / bs;
    }
};
struct System { std::vector<Table> tables; std::vector<CrossTableLookup> ctls; VF compress_challenges; };

/* permutation batches (permutation.rs:262-283): instances = pairs x challenges, chunked by batch_size;
 * instance i of a batch uses challenge_sets[i].challenges[chal]. */
struct PermInstance { const PermutationPair* pair; Challenge ch; };
inline std::vector<std::vector<PermInstance>> permutation_batches(const Table& t, const std::vector<std::vector<Challenge>>& sets, const Config& c) {
    std::vector<std::pair<const PermutationPair*, int>> all;
    for (auto& p : t.permutation_pairs)
        for (uint32_t ch = 0; ch < c.num_challenges; ch++) all.push_back({&p, (int)ch});
    std::vector<std::vector<PermInstance>> out;
    size_t bs = (size_t)t.permutation_batch_size();
    for (size_t s = 0; s < all.size(); s += bs) {
        std::vector<PermInstance> b;
        for (size_t i = 0; i < bs && s + i < all.size(); i++) b.push_back({all[s + i].first, sets[i][all[s + i].second]});
        out.push_back(b);
    }
    return out;
}
/* compute_permutation_z_poly (permutation.rs:129-160): exclusive prefix products of prod(lhs)/prod(rhs) */
inline VF permutation_z(const std::vector<PermInstance>& inst, const VF& trace, size_t n) {
    VF z(n);
    F acc = 1;
    for (size_t i = 0; i < n; i++) {
        z[i] = acc;
        F num = 1, den = 1;
        for (auto& in : inst) {
            F l = in.ch.gamma, r = in.ch.gamma, w = 1;
            for (auto& cp : in.pair->column_pairs) {
                l = gl_add(l, gl_mul(trace[(size_t)cp.first * n + i], w));
                r = gl_add(r, gl_mul(trace[(size_t)cp.second * n + i], w));
                w = gl_mul(w, in.ch.beta);
            }
            num = gl_mul(num, l);
            den = gl_mul(den, r);
        }
        acc = gl_mul(acc, gl_mul(num, gl_inv(den)));
    }
    return z;
}

/* ------------------------------------------------------------------ eval_vanishing_poly (vanishing_poly.rs:20-47) */
template <class O>
struct CtlVars { P<O> local_z, next_z; Challenge ch; const std::vector<Column>* columns; bool has_filter; const Column* filter; };

template <class O>
void eval_vanishing_poly(const Table& t, const Config& cfg, const P<O>* lv, const P<O>* nv, const std::vector<P<O>>* perm_local,
                         const std::vector<P<O>>* perm_next, const std::vector<std::vector<Challenge>>* perm_sets,
                         const std::vector<CtlVars<O>>& ctl, Consumer<O>& cons, const std::function<void(const P<O>*, const P<O>*, Consumer<O>&)>& eval) {
    typedef P<O> T;
    eval(lv, nv, cons);
    if (perm_local) { /* eval_permutation_checks (permutation.rs:302-360) */
        for (auto& z : *perm_local) cons.constraint_first_row(z - T::one());
        auto batches = permutation_batches(t, *perm_sets, cfg);
        for (size_t i = 0; i < batches.size(); i++) {
            T lhs = T::one(), rhs = T::one();
            for (auto& in : batches[i]) {
                T l = T::zero(), r = T::zero();
                auto& cp = in.pair->column_pairs;
                for (size_t k = cp.size(); k-- > 0;) { /* ReducingFactor::reduce_ext */
                    l = l * in.ch.beta + lv[cp[k].first];
                    r = r * in.ch.beta + lv[cp[k].second];
                }
                lhs = lhs * (l + T::c(in.ch.gamma));
                rhs = rhs * (r + T::c(in.ch.gamma));
            }
            cons.constraint((*perm_next)[i] * rhs - (*perm_local)[i] * lhs);
        }
    }
    for (auto& cv : ctl) { /* eval_cross_table_lookup_checks (cross_table_lookup.rs:380-419) */
        auto comb = [&](const T* v) {
            std::vector<T> ev;
            for (auto& c : *cv.columns) ev.push_back(c.template eval<O>(v));
            return combine<O>(cv.ch, ev);
        };
        auto filt = [&](const T* v) { return cv.has_filter ? cv.filter->template eval<O>(v) : T::one(); };
        T lf = filt(lv), nf = filt(nv);
        auto select = [&](T f, T x) { return f * x + T::one() - f; };
        cons.constraint_first_row(cv.local_z - select(lf, comb(lv)));
        cons.constraint_transition(cv.next_z - cv.local_z * select(nf, comb(nv)));
    }
}

/* ------------------------------------------------------------------ prove_single_table (prover.rs:330-567) */
inline std::string prove_single_table(const Table& t, const Config& cfg, const VF& trace, size_t n, const Batch& trace_commit,
                                      const std::vector<CtlZ>& ctl_zs, Challenger& ch, StarkProof& out) {
    uint32_t degree_bits = orc_log2_strict(n);
    FriParams fp = fri_params(cfg, degree_bits);
    if (fp.total_arities() > degree_bits + cfg.rate_bits - cfg.cap_height) return "FRI total reduction arity is too large.";
    ch.compact();
    std::vector<std::vector<Challenge>> perm_sets;
    std::vector<VF> z_polys;
    size_t num_perm_zs = 0;
    if (t.uses_permutation_args()) {
        for (int s = 0; s < t.permutation_batch_size(); s++) { /* get_n_grand_product_challenge_sets */
            std::vector<Challenge> set;
            for (uint32_t k = 0; k < cfg.num_challenges; k++) { F b = ch.get_challenge(); F g = ch.get_challenge(); set.push_back({b, g}); }
            perm_sets.push_back(set);
        }
        for (auto& b : permutation_batches(t, perm_sets, cfg)) z_polys.push_back(permutation_z(b, trace, n));
        num_perm_zs = z_polys.size();
    }
    for (auto& cz : ctl_zs) z_polys.push_back(cz.z);
    if (z_polys.empty()) return "No CTL?";
    VF zflat(z_polys.size() * n);
    for (size_t i = 0; i < z_polys.size(); i++) memcpy(&zflat[i * n], z_polys[i].data(), n * 8);
    Batch zs_commit = commit(zflat, z_polys.size(), n, false, cfg);
    ch.observe_cap(zs_commit.cap);
    VF alphas = ch.get_n(cfg.num_challenges);

    /* compute_quotient_polys (prover.rs:571-705) */
    int qdf = t.quotient_degree_factor();
    uint32_t qdb = 0;
    while ((1 << qdb) < qdf) qdb++;
    if (qdb > cfg.rate_bits) return "Having constraints of degree higher than the rate is not supported yet.";
    size_t step = (size_t)1 << (cfg.rate_bits - qdb), next_step = (size_t)1 << qdb;
    size_t size = n << qdb;
    uint32_t lde_bits = degree_bits + cfg.rate_bits;
    VF sel0(n, 0), sell(n, 0), lag_first(size), lag_last(size);
    sel0[0] = 1;
    sell[n - 1] = 1;
    orc_interpolate_poly(sel0.data(), n);
    orc_interpolate_poly(sell.data(), n);
    orc_evaluate_poly_with_offset(sel0.data(), n, GL_GEN, next_step, lag_first.data()); /* lde_onto_coset(qdb) */
    orc_evaluate_poly_with_offset(sell.data(), n, GL_GEN, next_step, lag_last.data());
    F g_pow_n = gl_pow(GL_GEN, n); /* ZeroPolyOnCoset::new (zero_poly_coset.rs:19-33) */
    VF zh_inv(next_step);
    {
        F w = gl_root_of_unity((int)qdb), x = 1;
        for (size_t i = 0; i < next_step; i++) { zh_inv[i] = gl_inv(gl_sub(gl_mul(g_pow_n, x), 1)); x = gl_mul(x, w); }
    }
    F last = gl_inv(gl_root_of_unity((int)degree_bits));
    F gq = gl_root_of_unity((int)(degree_bits + qdb));
    std::vector<VF> qvals(cfg.num_challenges, VF(size));
    std::vector<F> coset(size);
    { F x = GL_GEN; for (size_t i = 0; i < size; i++) { coset[i] = x; x = gl_mul(x, gq); } }
    auto batches_perm = t.uses_permutation_args() ? permutation_batches(t, perm_sets, cfg) : std::vector<std::vector<PermInstance>>();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; i++) {
        size_t i_next = (i + next_step) % size;
        size_t li = orc_bitrev(i * step, lde_bits), ln = orc_bitrev(i_next * step, lde_bits);
        std::vector<P<FOps>> lv(t.columns), nv(t.columns);
        for (int c = 0; c < t.columns; c++) { lv[c] = P<FOps>(trace_commit.leaf(li)[c]); nv[c] = P<FOps>(trace_commit.leaf(ln)[c]); }
        Consumer<FOps> cons(alphas, P<FOps>(gl_sub(coset[i], last)), P<FOps>(lag_first[i]), P<FOps>(lag_last[i]));
        std::vector<P<FOps>> pl, pn;
        for (size_t k = 0; k < num_perm_zs; k++) { pl.push_back(P<FOps>(zs_commit.leaf(li)[k])); pn.push_back(P<FOps>(zs_commit.leaf(ln)[k])); }
        std::vector<CtlVars<FOps>> cv;
        for (size_t k = 0; k < ctl_zs.size(); k++)
            cv.push_back({P<FOps>(zs_commit.leaf(li)[num_perm_zs + k]), P<FOps>(zs_commit.leaf(ln)[num_perm_zs + k]), ctl_zs[k].ch,
                          &ctl_zs[k].columns, ctl_zs[k].has_filter, &ctl_zs[k].filter});
        eval_vanishing_poly<FOps>(t, cfg, lv.data(), nv.data(), num_perm_zs ? &pl : nullptr, num_perm_zs ? &pn : nullptr, &perm_sets, cv, cons,
                                  t.eval_base);
        for (uint32_t j = 0; j < cfg.num_challenges; j++) qvals[j][i] = gl_mul(cons.accs[j].v, zh_inv[i % next_step]);
    }
    /* coset_ifft, trim_to_len(n*qdf), chunks(n)  (prover.rs:700-704, :463-478) */
    VF chunks;
    size_t nchunks = 0;
    for (uint32_t j = 0; j < cfg.num_challenges; j++) {
        orc_interpolate_poly_with_offset(qvals[j].data(), size, GL_GEN);
        if (cfg.check_quotient_degree)
            for (size_t k = n * qdf; k < size; k++)
                if (qvals[j][k] != 0) return "Quotient has failed, the vanishing polynomial is not divisible by Z_H";
        for (int c = 0; c < qdf; c++) { chunks.insert(chunks.end(), qvals[j].begin() + (size_t)c * n, qvals[j].begin() + (size_t)(c + 1) * n); nchunks++; }
    }
    Batch q_commit = commit(chunks, nchunks, n, true, cfg);
    ch.observe_cap(q_commit.cap);
    E zeta = ch.get_ext();
    F g = gl_root_of_unity((int)degree_bits);
    {
        E zp = zeta;
        for (uint32_t k = 0; k < degree_bits; k++) zp = gl2_mul(zp, zp);
        if (gl2_eq(zp, e_one())) return "Opening point is in the subgroup.";
    }
    /* StarkOpeningSet::new (proof.rs:199-246) */
    E zeta_next = gl2_scalar_mul(zeta, g);
    auto evalc = [&](const Batch& b, E z) { VE r; for (size_t c = 0; c < b.ncols; c++) r.push_back(poly_eval_ext_base(&b.coeffs[c * n], n, z)); return r; };
    OpeningSet os;
    os.local_values = evalc(trace_commit, zeta);
    os.next_values = evalc(trace_commit, zeta_next);
    os.zs = evalc(zs_commit, zeta);
    os.zs_next = evalc(zs_commit, zeta_next);
    F g_inv = gl_inv(g);
    for (size_t c = num_perm_zs; c < zs_commit.ncols; c++) os.ctl_zs_last.push_back(orc_poly_eval(&zs_commit.coeffs[c * n], n, g_inv));
    os.quotient = evalc(q_commit, zeta);
    /* observe_openings(to_fri_openings) (proof.rs:248-283, challenges.rs:15-23) */
    for (auto& v : os.local_values) ch.observe_ext(v);
    for (auto& v : os.zs) ch.observe_ext(v);
    for (auto& v : os.quotient) ch.observe_ext(v);
    for (auto& v : os.next_values) ch.observe_ext(v);
    for (auto& v : os.zs_next) ch.observe_ext(v);
    for (auto& v : os.ctl_zs_last) ch.observe_ext(e_from(v));
    /* fri_instance (stark.rs:87-150) */
    FriInstance inst;
    inst.oracle_num_polys = {(size_t)t.columns, zs_commit.ncols, q_commit.ncols};
    FriBatch b0, b1, b2;
    b0.point = zeta;
    for (int c = 0; c < t.columns; c++) b0.polys.push_back({0, c});
    for (size_t c = 0; c < zs_commit.ncols; c++) b0.polys.push_back({1, (int)c});
    for (size_t c = 0; c < q_commit.ncols; c++) b0.polys.push_back({2, (int)c});
    b1.point = zeta_next;
    for (int c = 0; c < t.columns; c++) b1.polys.push_back({0, c});
    for (size_t c = 0; c < zs_commit.ncols; c++) b1.polys.push_back({1, (int)c});
    b2.point = e_from(g_inv);
    for (size_t c = num_perm_zs; c < zs_commit.ncols; c++) b2.polys.push_back({1, (int)c});
    inst.batches = {b0, b1, b2};
    out.trace_cap = trace_commit.cap;
    out.zs_cap = zs_commit.cap;
    out.quotient_cap = q_commit.cap;
    out.openings = os;
    out.fri = prove_openings(inst, {&trace_commit, &zs_commit, &q_commit}, ch, fp, cfg);
    return "";
}

/* ------------------------------------------------------------------ prove_with_traces (prover.rs:79-327) */
struct AllProof { std::vector<StarkProof> proofs; VF compress_challenges; };

inline std::string prove_with_traces(const System& sys, const Config& cfg, const std::vector<VF>& traces, const std::vector<size_t>& ns,
                                     AllProof& out) {
    size_t T = sys.tables.size();
    std::vector<Batch> commits;
    for (size_t i = 0; i < T; i++) commits.push_back(commit(traces[i], sys.tables[i].columns, ns[i], false, cfg));
    Challenger ch;
    for (auto& c : commits) ch.observe_cap(c.cap);
    /* cross_table_lookup_data (cross_table_lookup.rs:224-282) */
    std::vector<Challenge> ctl_ch;
    for (uint32_t k = 0; k < cfg.num_challenges; k++) { F b = ch.get_challenge(); F g = ch.get_challenge(); ctl_ch.push_back({b, g}); }
    std::vector<std::vector<CtlZ>> per_table(T);
    for (auto& ctl : sys.ctls)
        for (auto& c : ctl_ch) {
            for (auto& lt : ctl.looking) {
                CtlZ z{{}, c, lt.columns, lt.has_filter, lt.filter};
                std::string e = partial_products(traces[lt.table], ns[lt.table], lt, c, z.z);
                if (!e.empty()) return e;
                per_table[lt.table].push_back(std::move(z));
            }
            if (!ctl.has_looked) continue;
            CtlZ z{{}, c, ctl.looked.columns, ctl.looked.has_filter, ctl.looked.filter};
            std::string e = partial_products(traces[ctl.looked.table], ns[ctl.looked.table], ctl.looked, c, z.z);
            if (!e.empty()) return e;
            per_table[ctl.looked.table].push_back(std::move(z));
        }
    out.proofs.resize(T);
    for (size_t i = 0; i < T; i++) {
        std::string e = prove_single_table(sys.tables[i], cfg, traces[i], ns[i], commits[i], per_table[i], ch, out.proofs[i]);
        if (!e.empty()) return sys.tables[i].name + ": " + e;
    }
    out.compress_challenges = sys.compress_challenges;
    return "";
}

/* ------------------------------------------------------------------ wire format (serialization.rs:349-393) */
struct Writer {
    std::vector<uint8_t> buf;
    void u8(uint8_t x) { buf.push_back(x); }
    void u32(uint32_t x) { for (int i = 0; i < 4; i++) buf.push_back((uint8_t)(x >> (8 * i))); }
    void field(F x) { x = gl_canon(x); for (int i = 0; i < 8; i++) buf.push_back((uint8_t)(x >> (8 * i))); }
    void ext(E x) { field(x.c0); field(x.c1); }
    void field_vec(const VF& v) { u32((uint32_t)v.size()); for (F x : v) field(x); }
    void ext_vec(const VE& v) { u32((uint32_t)v.size()); for (E x : v) ext(x); }
    void raw64(uint64_t x) { for (int i = 0; i < 8; i++) buf.push_back((uint8_t)(x >> (8 * i))); }
    /* write_hash = h.to_bytes() (serialization.rs:115-117): HashOut -> canonical u64s; BytesHash<32> -> its 32 bytes */
    void hash(const Hash& h) { for (int i = 0; i < 4; i++) { if (orc_get_hasher() == 1) raw64(h.e[i]); else field(h.e[i]); } }
    void cap(const Cap& c) { u32((uint32_t)c.size()); for (auto& h : c) hash(h); }
    void merkle_proof(const std::vector<Hash>& s) { u8((uint8_t)s.size()); for (auto& h : s) hash(h); }
    void proof(const StarkProof& p) {
        cap(p.trace_cap); cap(p.zs_cap); cap(p.quotient_cap);
        ext_vec(p.openings.local_values); ext_vec(p.openings.next_values); ext_vec(p.openings.zs); ext_vec(p.openings.zs_next);
        field_vec(p.openings.ctl_zs_last); ext_vec(p.openings.quotient);
        u32((uint32_t)p.fri.commit_caps.size());
        for (auto& c : p.fri.commit_caps) cap(c);
        u32((uint32_t)p.fri.rounds.size());
        for (auto& r : p.fri.rounds) {
            u32((uint32_t)r.initial.size());
            for (auto& ip : r.initial) { field_vec(ip.first); merkle_proof(ip.second); }
            u32((uint32_t)r.steps.size());
            for (auto& s : r.steps) { ext_vec(s.evals); merkle_proof(s.siblings); }
        }
        ext_vec(p.fri.final_poly);
        field(p.fri.pow_witness);
    }
    void all(const AllProof& a) { u32((uint32_t)a.proofs.size()); for (auto& p : a.proofs) proof(p); field_vec(a.compress_challenges); }
};
struct Reader {
    const uint8_t* p; size_t n, pos = 0; bool ok = true;
    Reader(const uint8_t* d, size_t len) : p(d), n(len) {}
    uint8_t u8() { if (pos + 1 > n) { ok = false; return 0; } return p[pos++]; }
    uint32_t u32() { if (pos + 4 > n) { ok = false; return 0; } uint32_t x = 0; for (int i = 0; i < 4; i++) x |= (uint32_t)p[pos + i] << (8 * i); pos += 4; return x; }
    F field() { if (pos + 8 > n) { ok = false; return 0; } F x = 0; for (int i = 0; i < 8; i++) x |= (F)p[pos + i] << (8 * i); pos += 8; return x; }
    E ext() { F a = field(); F b = field(); return gl2_make(a, b); }
    VF field_vec() { uint32_t k = u32(); VF v; for (uint32_t i = 0; i < k && ok; i++) v.push_back(field()); return v; }
    VE ext_vec() { uint32_t k = u32(); VE v; for (uint32_t i = 0; i < k && ok; i++) v.push_back(ext()); return v; }
    Hash hash() { Hash h; for (int i = 0; i < 4; i++) h.e[i] = field(); return h; }
    Cap cap() { uint32_t k = u32(); Cap c; for (uint32_t i = 0; i < k && ok; i++) c.push_back(hash()); return c; }
    std::vector<Hash> merkle_proof() { uint8_t k = u8(); std::vector<Hash> s; for (int i = 0; i < k && ok; i++) s.push_back(hash()); return s; }
    StarkProof proof() {
        StarkProof q;
        q.trace_cap = cap(); q.zs_cap = cap(); q.quotient_cap = cap();
        q.openings.local_values = ext_vec(); q.openings.next_values = ext_vec(); q.openings.zs = ext_vec(); q.openings.zs_next = ext_vec();
        q.openings.ctl_zs_last = field_vec(); q.openings.quotient = ext_vec();
        uint32_t nc = u32();
        for (uint32_t i = 0; i < nc && ok; i++) q.fri.commit_caps.push_back(cap());
        uint32_t nr = u32();
        for (uint32_t i = 0; i < nr && ok; i++) {
            FriQueryRound r;
            uint32_t ni = u32();
            for (uint32_t k = 0; k < ni && ok; k++) { VF v = field_vec(); auto s = merkle_proof(); r.initial.push_back({v, s}); }
            uint32_t ns = u32();
            for (uint32_t k = 0; k < ns && ok; k++) { FriQueryStep s; s.evals = ext_vec(); s.siblings = merkle_proof(); r.steps.push_back(s); }
            q.fri.rounds.push_back(r);
        }
        q.fri.final_poly = ext_vec();
        q.fri.pow_witness = field();
        return q;
    }
    AllProof all() { AllProof a; uint32_t k = u32(); for (uint32_t i = 0; i < k && ok; i++) a.proofs.push_back(proof()); a.compress_challenges = field_vec(); return a; }
};

/* ------------------------------------------------------------------ verify_proof (verifier.rs, get_challenges.rs) */
inline void eval_l_0_and_l_last(uint32_t log_n, E x, E& l0, E& ll) { /* verifier.rs:380-396 */
    size_t n = (size_t)1 << log_n;
    F g = gl_root_of_unity((int)log_n);
    E zx = x;
    for (uint32_t k = 0; k < log_n; k++) zx = gl2_mul(zx, zx);
    zx = gl2_sub(zx, e_one());
    E inv0 = gl2_inv(gl2_scalar_mul(gl2_sub(x, e_one()), n % GL_P));
    E invl = gl2_inv(gl2_scalar_mul(gl2_sub(gl2_scalar_mul(x, g), e_one()), n % GL_P));
    l0 = gl2_mul(zx, inv0);
    ll = gl2_mul(zx, invl);
}

inline std::string verify_all(const System& sys, const Config& cfg, const AllProof& ap) {
    size_t T = sys.tables.size();
    if (ap.proofs.size() != T) return "wrong number of proofs";
    Challenger ch;
    for (auto& p : ap.proofs) ch.observe_cap(p.trace_cap);
    std::vector<Challenge> ctl_ch;
    for (uint32_t k = 0; k < cfg.num_challenges; k++) { F b = ch.get_challenge(); F g = ch.get_challenge(); ctl_ch.push_back({b, g}); }
    /* CtlCheckVars::from_proofs (cross_table_lookup.rs:337-377): consume each table's ctl zs in registry order */
    std::vector<size_t> cursor(T, 0);
    std::vector<std::vector<CtlVars<EOps>>> ctl_vars(T);
    std::vector<size_t> nperm(T);
    for (size_t i = 0; i < T; i++) nperm[i] = sys.tables[i].num_permutation_batches(cfg);
    for (auto& ctl : sys.ctls)
        for (auto& c : ctl_ch) {
            auto push = [&](const TableWithColumns& tw) {
                const StarkProof& p = ap.proofs[tw.table];
                size_t k = nperm[tw.table] + cursor[tw.table]++;
                if (k >= p.openings.zs.size() || k >= p.openings.zs_next.size()) return false;
                ctl_vars[tw.table].push_back({P<EOps>(p.openings.zs[k]), P<EOps>(p.openings.zs_next[k]), c, &tw.columns, tw.has_filter, &tw.filter});
                return true;
            };
            for (auto& lt : ctl.looking) if (!push(lt)) return "ctl shape";
            if (ctl.has_looked && !push(ctl.looked)) return "ctl shape";
        }
    for (size_t i = 0; i < T; i++) {
        const Table& t = sys.tables[i];
        const StarkProof& p = ap.proofs[i];
        ch.compact();
        /* recover_degree_bits (proof.rs): initial merkle proof length + cap_height - rate_bits */
        if (p.fri.rounds.empty() || p.fri.rounds[0].initial.empty()) return "empty proof";
        uint32_t lde_bits = (uint32_t)p.fri.rounds[0].initial[0].second.size() + cfg.cap_height;
        uint32_t degree_bits = lde_bits - cfg.rate_bits;
        std::vector<std::vector<Challenge>> perm_sets;
        if (t.uses_permutation_args())
            for (int s = 0; s < t.permutation_batch_size(); s++) {
                std::vector<Challenge> set;
                for (uint32_t k = 0; k < cfg.num_challenges; k++) { F b = ch.get_challenge(); F g = ch.get_challenge(); set.push_back({b, g}); }
                perm_sets.push_back(set);
            }
        ch.observe_cap(p.zs_cap);
        VF alphas = ch.get_n(cfg.num_challenges);
        ch.observe_cap(p.quotient_cap);
        E zeta = ch.get_ext();
        const OpeningSet& os = p.openings;
        for (auto& v : os.local_values) ch.observe_ext(v);
        for (auto& v : os.zs) ch.observe_ext(v);
        for (auto& v : os.quotient) ch.observe_ext(v);
        for (auto& v : os.next_values) ch.observe_ext(v);
        for (auto& v : os.zs_next) ch.observe_ext(v);
        for (auto& v : os.ctl_zs_last) ch.observe_ext(e_from(v));
        FriChallenges fc = fri_challenges(ch, p.fri, degree_bits, cfg);
        /* validate_proof_shape (verifier.rs:310-360) */
        size_t num_zs = nperm[i] + ctl_vars[i].size();
        size_t ncap = (size_t)1 << cfg.cap_height;
        if (p.trace_cap.size() != ncap || p.zs_cap.size() != ncap || p.quotient_cap.size() != ncap) return t.name + ": cap shape";
        if (os.local_values.size() != (size_t)t.columns || os.next_values.size() != (size_t)t.columns || os.zs.size() != num_zs ||
            os.zs_next.size() != num_zs || os.ctl_zs_last.size() != ctl_vars[i].size() ||
            os.quotient.size() != (size_t)t.quotient_degree_factor() * cfg.num_challenges)
            return t.name + ": opening set shape";
        /* verify_stark_proof_with_challenges (verifier.rs:214-300) */
        E l0, ll;
        eval_l_0_and_l_last(degree_bits, zeta, l0, ll);
        F last = gl_inv(gl_root_of_unity((int)degree_bits));
        Consumer<EOps> cons(alphas, P<EOps>(gl2_sub(zeta, e_from(last))), P<EOps>(l0), P<EOps>(ll));
        std::vector<P<EOps>> lv, nv, pl, pn;
        for (auto& v : os.local_values) lv.push_back(P<EOps>(v));
        for (auto& v : os.next_values) nv.push_back(P<EOps>(v));
        for (size_t k = 0; k < nperm[i]; k++) { pl.push_back(P<EOps>(os.zs[k])); pn.push_back(P<EOps>(os.zs_next[k])); }
        eval_vanishing_poly<EOps>(t, cfg, lv.data(), nv.data(), nperm[i] ? &pl : nullptr, nperm[i] ? &pn : nullptr, &perm_sets, ctl_vars[i], cons,
                                  t.eval_ext);
        E zpow = zeta;
        for (uint32_t k = 0; k < degree_bits; k++) zpow = gl2_mul(zpow, zpow);
        E zh = gl2_sub(zpow, e_one());
        int qdf = t.quotient_degree_factor();
        for (uint32_t j = 0; j < cfg.num_challenges; j++) {
            E s = e_zero();
            for (int k = qdf; k-- > 0;) s = gl2_add(gl2_mul(s, zpow), os.quotient[(size_t)j * qdf + k]);
            if (!gl2_eq(cons.accs[j].v, gl2_mul(zh, s))) return "Mismatch between evaluation and opening of quotient polynomial in " + t.name;
        }
        FriInstance inst;
        inst.oracle_num_polys = {(size_t)t.columns, num_zs, (size_t)qdf * cfg.num_challenges};
        F g = gl_root_of_unity((int)degree_bits);
        FriBatch b0, b1, b2;
        b0.point = zeta;
        for (int c = 0; c < t.columns; c++) b0.polys.push_back({0, c});
        for (size_t c = 0; c < num_zs; c++) b0.polys.push_back({1, (int)c});
        for (size_t c = 0; c < (size_t)qdf * cfg.num_challenges; c++) b0.polys.push_back({2, (int)c});
        b1.point = gl2_scalar_mul(zeta, g);
        for (int c = 0; c < t.columns; c++) b1.polys.push_back({0, c});
        for (size_t c = 0; c < num_zs; c++) b1.polys.push_back({1, (int)c});
        b2.point = e_from(gl_inv(g));
        for (size_t c = nperm[i]; c < num_zs; c++) b2.polys.push_back({1, (int)c});
        inst.batches = {b0, b1, b2};
        std::vector<VE> openings(3);
        openings[0] = os.local_values; openings[0].insert(openings[0].end(), os.zs.begin(), os.zs.end()); openings[0].insert(openings[0].end(), os.quotient.begin(), os.quotient.end());
        openings[1] = os.next_values; openings[1].insert(openings[1].end(), os.zs_next.begin(), os.zs_next.end());
        for (F v : os.ctl_zs_last) openings[2].push_back(e_from(v));
        std::string e = verify_fri(inst, openings, fc, {p.trace_cap, p.zs_cap, p.quotient_cap}, p.fri, fri_params(cfg, degree_bits), cfg);
        if (!e.empty()) return t.name + ": " + e;
    }
    /* verify_cross_table_lookups (cross_table_lookup.rs:560-600): prod(looking Z_last) == looked Z_last */
    std::vector<size_t> cur(T, 0);
    for (auto& ctl : sys.ctls)
        for (uint32_t k = 0; k < cfg.num_challenges; k++) {
            F prod = 1;
            for (auto& lt : ctl.looking) prod = gl_mul(prod, ap.proofs[lt.table].openings.ctl_zs_last[cur[lt.table]++]);
            if (!ctl.has_looked) continue;
            F looked = ap.proofs[ctl.looked.table].openings.ctl_zs_last[cur[ctl.looked.table]++];
            if (!ctl.complete) continue; /* partial CTL (a side's table is outside this system): nothing to compare */
            if (prod != looked) return "Cross-table lookup verification failed.";
        }
    return "";
}

}  // namespace orc
#endif
