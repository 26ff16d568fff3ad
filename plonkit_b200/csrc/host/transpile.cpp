// Host-side circuit intake for the prove path (plain C++, no CUDA): the iden3 `.r1cs` parser and the
// R1CS -> width-4 gate transpilation, i.e. the compiled counterpart of plonkit_b200/circuit.py::_transpile and
// reader.py::load_r1cs_from_bin, which stay the readable statement of the same layout (the tests hold the two equal).
//
// Reference: src/r1cs_file.rs:100-154 + src/reader.rs:227-241 (file format), src/circom_circuit.rs:75-133 (variable
// allocation, skipped `0 * LC = 0`), src/transpile.rs:92-139 (the wrapper around bellman's adaptor, which is not in the
// reference tree: strict mode accepts exactly the constraint shapes the reference's golden vectors pin; everything else is
// this repository's own layout, byte parity unpinned — DESIGN.md section 8).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include <utility>
#include <vector>

#include "fr_host.hpp"

using namespace phost;

namespace {

thread_local std::string g_error;

struct R1cs {
    uint64_t num_inputs = 0, num_aux = 0, num_variables = 0;
    std::vector<uint64_t> off;        // 3 * num_constraints + 1 offsets: A, B, C of constraint 0, A of constraint 1, ...
    std::vector<uint32_t> var;
    std::vector<Fr> coef;             // canonical
    std::vector<uint64_t> wire_map;   // section 3 of the file (empty when built from arrays)
    uint64_t num_constraints() const { return off.empty() ? 0 : (off.size() - 1) / 3; }
};

struct Gates {
    std::vector<uint32_t> wire[4];
    std::vector<Fr> sel[7];           // q_a, q_b, q_c, q_d, q_m, q_const, q_dnext
    uint64_t num_vars = 0, num_direct = 0, hints = 0;
    std::vector<uint64_t> prog_off;   // per introduced variable: const + sum coef * earlier variable
    std::vector<uint32_t> prog_var;
    std::vector<Fr> prog_coef, prog_const;
    std::vector<uint32_t> stat_constraint, stat_gates;
    uint64_t detail[6] = {0, 0, 0, 0, 0, 0};  // error detail: code, constraint, aux, (constant limbs follow in err_const)
    Fr err_const = zero();
    uint64_t rows() const { return wire[0].size(); }
};

typedef std::pair<uint32_t, Fr> Term;

uint32_t rd32(const uint8_t* p) { uint32_t x; memcpy(&x, p, 4); return x; }
uint64_t rd64(const uint8_t* p) { uint64_t x; memcpy(&x, p, 8); return x; }

const uint8_t PRIME_LE[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                              0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};

bool fail(const char* msg) {
    g_error = msg;
    return false;
}

bool parse_bin(const uint8_t* buf, uint64_t len, R1cs& r) {
    if (len < 12 || memcmp(buf, "r1cs", 4) != 0) return fail("Invalid magic number");
    if (rd32(buf + 4) != 1) return fail("Unsupported version");
    const uint32_t num_sections = rd32(buf + 8);
    uint64_t pos = 12;
    uint64_t sec_off[4] = {0, 0, 0, 0}, sec_size[4] = {0, 0, 0, 0};
    bool have[4] = {false, false, false, false};
    for (uint32_t s = 0; s < num_sections; ++s) {
        if (pos + 12 > len) return fail("r1cs file is truncated");
        const uint32_t st = rd32(buf + pos);
        const uint64_t ss = rd64(buf + pos + 4);
        pos += 12;
        if (st >= 1 && st <= 3) { sec_off[st] = pos; sec_size[st] = ss; have[st] = true; }
        if (ss > len - pos) { pos = len; continue; }  // a section running off the end: whoever reads it finds out
        pos += ss;
    }
    if (!have[1]) return fail("r1cs file lacks a header, constraint or wire-map section");
    const uint64_t h = sec_off[1];
    if (h + 4 > len) return fail("r1cs file is truncated");
    const uint32_t field_size = rd32(buf + h);
    if (sec_size[1] != 32 + (uint64_t)field_size) return fail("Invalid header section size");
    if (field_size != 32) return fail("This parser only supports 32-byte fields");
    if (h + 64 > len) return fail("r1cs file is truncated");
    if (memcmp(buf + h + 4, PRIME_LE, 32) != 0) return fail("This parser only supports bn256");
    const uint32_t n_wires = rd32(buf + h + 36), n_pub_out = rd32(buf + h + 40), n_pub_in = rd32(buf + h + 44);
    const uint32_t n_constraints = rd32(buf + h + 60);
    if (!have[2] || !have[3]) return fail("r1cs file lacks a header, constraint or wire-map section");
    uint64_t p = sec_off[2];
    const uint64_t end = len;
    r.off.assign(1, 0);
    // the count comes from the file: never reserve more than the bytes that are left could hold (4 per linear combination)
    r.off.reserve(std::min<uint64_t>(3 * (uint64_t)n_constraints, (p <= len ? len - p : 0) / 4) + 1);
    for (uint64_t c = 0; c < 3 * (uint64_t)n_constraints; ++c) {
        if (p + 4 > end) return fail("r1cs file is truncated");
        const uint32_t k = rd32(buf + p);
        p += 4;
        if ((uint64_t)k * 36 > end - p) return fail("r1cs file is truncated");
        for (uint32_t t = 0; t < k; ++t) {
            Fr v;
            memcpy(v.v, buf + p + 4, 32);
            if (!canonical(v.v)) return fail("coefficient is not in the field");
            r.var.push_back(rd32(buf + p));
            r.coef.push_back(v);
            p += 36;
        }
        r.off.push_back(r.var.size());
    }
    if (sec_size[3] != (uint64_t)n_wires * 8) return fail("Invalid map section size");
    if (sec_off[3] > len || sec_size[3] > len - sec_off[3]) return fail("r1cs file is truncated");
    r.wire_map.resize(n_wires);
    for (uint32_t i = 0; i < n_wires; ++i) r.wire_map[i] = rd64(buf + sec_off[3] + 8 * (uint64_t)i);
    if (n_wires && r.wire_map[0] != 0) return fail("Wire 0 should always be mapped to 0");
    r.num_inputs = 1 + (uint64_t)n_pub_in + n_pub_out;
    if (r.num_inputs > n_wires) return fail("more public inputs and outputs than wires");
    r.num_aux = (uint64_t)n_wires - r.num_inputs;
    r.num_variables = n_wires;
    return true;
}

// -> terms sorted by variable, equal variables merged, zero coefficients dropped; wire 0 is the constant ONE
void norm_lc(const R1cs& r, uint64_t lo, uint64_t hi, std::vector<Term>& terms, Fr& constant) {
    terms.clear();
    constant = zero();
    for (uint64_t j = lo; j < hi; ++j) {
        if (r.var[j] == 0) constant = add(constant, r.coef[j]);
        else terms.push_back(Term(r.var[j], r.coef[j]));
    }
    std::stable_sort(terms.begin(), terms.end(), [](const Term& a, const Term& b) { return a.first < b.first; });
    size_t w = 0;
    for (size_t i = 0; i < terms.size();) {
        Fr acc = terms[i].second;
        size_t j = i + 1;
        while (j < terms.size() && terms[j].first == terms[i].first) acc = add(acc, terms[j++].second);
        if (!is_zero(acc)) terms[w++] = Term(terms[i].first, acc);
        i = j;
    }
    terms.resize(w);
}

struct Builder {
    Gates& g;
    const Fr M1 = neg(from_u64(1));
    explicit Builder(Gates& gates) : g(gates) {}

    void row(uint32_t a, uint32_t b, uint32_t c, uint32_t d, const Fr (&q)[7]) {
        g.wire[0].push_back(a); g.wire[1].push_back(b); g.wire[2].push_back(c); g.wire[3].push_back(d);
        for (int s = 0; s < 7; ++s) g.sel[s].push_back(q[s]);
    }
    // a fresh variable whose value is sum(terms) + constant
    uint32_t new_var(const Term* terms, size_t n, const Fr& constant) {
        for (size_t i = 0; i < n; ++i) { g.prog_var.push_back(terms[i].first); g.prog_coef.push_back(terms[i].second); }
        g.prog_const.push_back(constant);
        g.prog_off.push_back(g.prog_var.size());
        return (uint32_t)g.num_vars++;
    }
    // gates enforcing sum(terms) + constant - out = 0 (has_out = false: the sum itself is zero); more than four summands run
    // as a chain: every row adds up to three terms to the running sum in d and hands it to the next row's d
    void chain(const std::vector<Term>& terms, const Fr& constant, bool has_out, uint32_t out) {
        std::vector<Term> items(terms);
        if (has_out) items.push_back(Term(out, M1));
        const Fr Z = zero();
        if (items.size() <= 4) {
            uint32_t w[4] = {0, 0, 0, 0};
            Fr q[7] = {Z, Z, Z, Z, Z, constant, Z};
            for (size_t i = 0; i < items.size(); ++i) { w[i] = items[i].first; q[i] = items[i].second; }
            row(w[0], w[1], w[2], w[3], q);
            return;
        }
        uint32_t acc = new_var(items.data(), 4, constant);
        {
            Fr q[7] = {items[0].second, items[1].second, items[2].second, items[3].second, Z, constant, M1};
            row(items[0].first, items[1].first, items[2].first, items[3].first, q);
        }
        size_t pos = 4;
        while (pos < items.size()) {
            const size_t take = std::min<size_t>(3, items.size() - pos);
            uint32_t w[3] = {0, 0, 0};
            Fr c[3] = {Z, Z, Z};
            for (size_t i = 0; i < take; ++i) { w[i] = items[pos + i].first; c[i] = items[pos + i].second; }
            const bool more = pos + take < items.size();
            if (more) {
                Term sum[4];
                sum[0] = Term(acc, from_u64(1));
                for (size_t i = 0; i < take; ++i) sum[1 + i] = items[pos + i];
                const uint32_t nxt = new_var(sum, 1 + take, Z);
                Fr q[7] = {c[0], c[1], c[2], from_u64(1), Z, Z, M1};
                row(w[0], w[1], w[2], acc, q);
                acc = nxt;
            } else {
                Fr q[7] = {c[0], c[1], c[2], from_u64(1), Z, Z, Z};
                row(w[0], w[1], w[2], acc, q);
            }
            pos += take;
        }
    }
    // -> (variable, coefficient) with LC == coefficient * variable; a fresh variable (and its gates) unless the combination
    // already is a single variable
    Term collapse(const std::vector<Term>& terms, const Fr& constant) {
        if (terms.size() == 1 && is_zero(constant)) return terms[0];
        const uint32_t t = new_var(terms.data(), terms.size(), constant);
        const Fr Z = zero();
        if (terms.size() == 2) {  // the pinned layout: (a = v1, b = v2, c = t), q_c = -1
            Fr q[7] = {terms[0].second, terms[1].second, M1, Z, Z, constant, Z};
            row(terms[0].first, terms[1].first, t, 0, q);
        } else {
            chain(terms, constant, true, t);
        }
        return Term(t, from_u64(1));
    }
};

// returns 0, or 1 (A / B not single variables), 2 (C side not pinned), 3 (contradiction); g.detail / g.err_const say where
int transpile(const R1cs& r, bool strict, Gates& g) {
    Builder b(g);
    const Fr Z = zero();
    g.num_direct = g.num_vars = r.num_variables;
    g.prog_off.assign(1, 0);
    for (uint64_t i = 1; i < r.num_inputs; ++i) {
        Fr q[7] = {b.M1, Z, Z, Z, Z, Z, Z};
        b.row((uint32_t)i, 0, 0, 0, q);
    }
    std::vector<Term> ta, tb, tc, lin;
    Fr ka, kb, kc;
    const uint64_t nc = r.num_constraints();
    for (uint64_t ci = 0; ci < nc; ++ci) {
        const uint64_t a0 = r.off[3 * ci], b0 = r.off[3 * ci + 1], c0 = r.off[3 * ci + 2], c1 = r.off[3 * ci + 3];
        if ((a0 == b0 || b0 == c0) && c0 == c1) continue;  // 0 * LC = 0 is ignored (circom_circuit.rs:122-123)
        const uint64_t before = g.rows();
        norm_lc(r, a0, b0, ta, ka);
        norm_lc(r, b0, c0, tb, kb);
        norm_lc(r, c0, c1, tc, kc);
        const bool ab_single = ta.size() == 1 && is_zero(ka) && tb.size() == 1 && is_zero(kb);
        const bool pinned = ab_single && ((tc.size() == 1 && is_zero(kc)) || tc.size() == 2);
        if (!pinned && strict) {
            g.detail[0] = ab_single ? 2 : 1;
            g.detail[1] = ci;
            g.detail[2] = tc.size();
            return (int)g.detail[0];
        }
        if (ta.empty() || tb.empty()) {
            // a constant factor: k * LC_other - LC_C = 0 is linear
            const Fr k = ta.empty() ? ka : kb;
            const std::vector<Term>& to = ta.empty() ? tb : ta;
            const Fr ko = ta.empty() ? kb : ka;
            lin.clear();
            for (const Term& t : to) lin.push_back(Term(t.first, mul(k, t.second)));
            for (const Term& t : tc) lin.push_back(Term(t.first, neg(t.second)));
            std::stable_sort(lin.begin(), lin.end(), [](const Term& x, const Term& y) { return x.first < y.first; });
            size_t w = 0;
            for (size_t i = 0; i < lin.size();) {
                Fr acc = lin[i].second;
                size_t j = i + 1;
                while (j < lin.size() && lin[j].first == lin[i].first) acc = add(acc, lin[j++].second);
                if (!is_zero(acc)) lin[w++] = Term(lin[i].first, acc);
                i = j;
            }
            lin.resize(w);
            const Fr constant = sub(mul(k, ko), kc);
            if (lin.empty()) {
                if (!is_zero(constant)) {
                    g.detail[0] = 3;
                    g.detail[1] = ci;
                    g.err_const = constant;
                    return 3;
                }
            } else {
                b.chain(lin, constant, false, 0);
            }
        } else {
            const Term x = b.collapse(ta, ka), y = b.collapse(tb, kb);
            const Fr qm = mul(x.second, y.second);
            if (tc.empty()) {
                Fr q[7] = {Z, Z, Z, Z, qm, neg(kc), Z};
                b.row(x.first, y.first, 0, 0, q);
            } else {
                const Term z = b.collapse(tc, kc);
                Fr q[7] = {Z, Z, neg(z.second), Z, qm, Z, Z};
                b.row(x.first, y.first, z.first, 0, q);
            }
        }
        ++g.hints;
        g.stat_constraint.push_back((uint32_t)ci);
        g.stat_gates.push_back((uint32_t)(g.rows() - before));
    }
    return 0;
}

}  // namespace

extern "C" {

struct ph_r1cs;
struct ph_gates;

const char* ph_last_error(void) { return g_error.c_str(); }

// parse an iden3 .r1cs image (src/r1cs_file.rs:100-154); 0 on success, else ph_last_error() says why
int ph_r1cs_parse_bin(const uint8_t* buf, uint64_t len, ph_r1cs** out) {
    R1cs* r = nullptr;
    try {
        r = new R1cs();
        if (!parse_bin(buf, len, *r)) {
            delete r;
            return 1;
        }
    } catch (const std::exception& e) {  // nothing unwinds across the C boundary
        delete r;
        g_error = std::string("r1cs parser: ") + e.what();
        return 1;
    }
    *out = reinterpret_cast<ph_r1cs*>(r);
    return 0;
}
// the same object from arrays: lc_off has 3 * num_constraints + 1 entries (A, B, C of every constraint), coefficients canonical
int ph_r1cs_from_csr(uint64_t num_inputs, uint64_t num_aux, uint64_t num_variables, uint64_t num_constraints, const uint64_t* lc_off,
                     const uint32_t* lc_var, const uint64_t* lc_coef, ph_r1cs** out) {
    R1cs* r = nullptr;
    try {
        r = new R1cs();
        r->num_inputs = num_inputs; r->num_aux = num_aux; r->num_variables = num_variables;
        r->off.assign(lc_off, lc_off + 3 * num_constraints + 1);
        const uint64_t nt = r->off.back();
        r->var.assign(lc_var, lc_var + nt);
        r->coef.resize(nt);
        for (uint64_t i = 0; i < nt; ++i) {
            memcpy(r->coef[i].v, lc_coef + 4 * i, 32);
            if (!canonical(r->coef[i].v)) {
                delete r;
                g_error = "coefficient is not in the field";
                return 1;
            }
        }
    } catch (const std::exception& e) {
        delete r;
        g_error = std::string("r1cs from arrays: ") + e.what();
        return 1;
    }
    *out = reinterpret_cast<ph_r1cs*>(r);
    return 0;
}
void ph_r1cs_free(ph_r1cs* h) { delete reinterpret_cast<R1cs*>(h); }
// out: num_inputs, num_aux, num_variables, num_constraints, num_terms, wire-map length
void ph_r1cs_header(const ph_r1cs* h, uint64_t out[6]) {
    const R1cs* r = reinterpret_cast<const R1cs*>(h);
    out[0] = r->num_inputs; out[1] = r->num_aux; out[2] = r->num_variables; out[3] = r->num_constraints();
    out[4] = r->var.size(); out[5] = r->wire_map.size();
}
void ph_r1cs_export(const ph_r1cs* h, uint64_t* lc_off, uint32_t* lc_var, uint64_t* lc_coef, uint64_t* wire_map) {
    const R1cs* r = reinterpret_cast<const R1cs*>(h);
    memcpy(lc_off, r->off.data(), r->off.size() * 8);
    if (!r->var.empty()) {
        memcpy(lc_var, r->var.data(), r->var.size() * 4);
        memcpy(lc_coef, r->coef.data(), r->coef.size() * 32);
    }
    if (wire_map && !r->wire_map.empty()) memcpy(wire_map, r->wire_map.data(), r->wire_map.size() * 8);
}

// R1CS -> width-4 gates.  0 on success; 1 / 2 = a constraint shape strict mode does not accept, 3 = a contradiction:
// detail = {code, constraint index, number of C-side variables, constant limbs 0..3}
int ph_transpile(const ph_r1cs* h, int strict, ph_gates** out, uint64_t detail[7]) {
    Gates* g = nullptr;
    int rc;
    try {
        g = new Gates();
        rc = transpile(*reinterpret_cast<const R1cs*>(h), strict != 0, *g);
    } catch (const std::exception& e) {
        delete g;
        g_error = std::string("transpiler: ") + e.what();
        return 4;
    }
    if (rc) {
        detail[0] = g->detail[0]; detail[1] = g->detail[1]; detail[2] = g->detail[2];
        memcpy(detail + 3, g->err_const.v, 32);
        delete g;
        return rc;
    }
    *out = reinterpret_cast<ph_gates*>(g);
    return 0;
}
void ph_gates_free(ph_gates* h) { delete reinterpret_cast<Gates*>(h); }
// out: rows, variables, direct variables, hints, constraints that produced gates, program terms
void ph_gates_header(const ph_gates* h, uint64_t out[6]) {
    const Gates* g = reinterpret_cast<const Gates*>(h);
    out[0] = g->rows(); out[1] = g->num_vars; out[2] = g->num_direct; out[3] = g->hints; out[4] = g->stat_gates.size();
    out[5] = g->prog_var.size();
}
// wire_idx: (4, n) uint32 and selectors: (7, n, 4) uint64, both zero-filled by the caller, n >= rows; program arrays sized
// from ph_gates_header (prog_off: variables - direct + 1); stats: one (constraint, gates) pair per counted constraint
void ph_gates_export(const ph_gates* h, uint64_t n, uint32_t* wire_idx, uint64_t* selectors, uint64_t* prog_off, uint32_t* prog_var,
                     uint64_t* prog_coef, uint64_t* prog_const, uint32_t* stat_constraint, uint32_t* stat_gates) {
    const Gates* g = reinterpret_cast<const Gates*>(h);
    const uint64_t rows = g->rows();
    for (int c = 0; c < 4; ++c)
        if (rows) memcpy(wire_idx + c * n, g->wire[c].data(), rows * 4);
    for (int s = 0; s < 7; ++s)
        if (rows) memcpy(selectors + (uint64_t)s * n * 4, g->sel[s].data(), rows * 32);
    memcpy(prog_off, g->prog_off.data(), g->prog_off.size() * 8);
    if (!g->prog_var.empty()) {
        memcpy(prog_var, g->prog_var.data(), g->prog_var.size() * 4);
        memcpy(prog_coef, g->prog_coef.data(), g->prog_coef.size() * 32);
    }
    if (!g->prog_const.empty()) memcpy(prog_const, g->prog_const.data(), g->prog_const.size() * 32);
    if (!g->stat_gates.empty()) {
        memcpy(stat_constraint, g->stat_constraint.data(), g->stat_constraint.size() * 4);
        memcpy(stat_gates, g->stat_gates.data(), g->stat_gates.size() * 4);
    }
}

}  // extern "C"
