// Native mmCIF ingest (north_star: "CIF parsing, unit-id construction ... stay on the host"): CIF texts ->
// the packed nucleotide records of an NtBatch (include/fr3d_b200.h, fr3d_nts_t layout) + identity columns, with a
// thread per structure.  Host code only (no CUDA): compiled into libfr3d_b200.so next to the kernels so that the
// records can be written straight into pinned memory.
//
// It follows fr3d_python_b200/cif/reader.py + packing.pack_specs step by step (those are the readable statement of
// the same logic and the parity oracle of tests/test_cif_ingest.py), which in turn restate the reference:
//   operators / assemblies, one atom per (operator, atom_site row)      fr3d/cif/reader.py:114-237, 486-629
//   residues = stable sort + group by (model, chain, component id, number, insertion code, symmetry)  :440-451
//   alternate conformers (common atoms copied into each)                :420-438
//   atom slots, duplicate averaging (AtomProxy), observed amino hydrogens   fr3d/data/base.py:150-166, components.py:405-445
//   previous O3' (map_unit_id_to_previous_O3)                           NA_pairwise_interactions.py:480-525
#include <algorithm>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/fr3d_b200.h"

namespace {

using sv = std::string_view;

thread_local char g_ingest_err[512] = "";
int ifail(const char *msg) {
    snprintf(g_ingest_err, sizeof(g_ingest_err), "%s", msg);
    return FR3D_E_ARG;
}

// ---------------------------------------------------------------- slot tables (from packing._slot_maps) ----
// atom names (at most 8 bytes in every table we ship; checked when the tables are loaded) packed into one integer
inline bool pack_name(std::string_view s, uint64_t &key) {
    if (s.size() > 8) return false;
    key = 0;
    memcpy(&key, s.data(), s.size());
    return true;
}

struct SlotMap {                       // ~ 25 entries: a linear scan over packed names beats hashing strings
    std::vector<uint64_t> keys;
    std::vector<int> slots;
    void set(uint64_t k, int slot) {
        for (size_t q = 0; q < keys.size(); ++q)
            if (keys[q] == k) { slots[q] = slot; return; }
        keys.push_back(k);
        slots.push_back(slot);
    }
    int find(uint64_t k) const {
        for (size_t q = 0; q < keys.size(); ++q)
            if (keys[q] == k) return slots[q];
        return -1;
    }
};

struct SeqInfo {
    int pcode = -1;
    int flags = 0;
    uint32_t dup = 0;
    SlotMap base, bb;
};
std::unordered_map<std::string, SeqInfo> g_seq;
struct Amino { int heavy, h1, h2; };
std::map<int, Amino> g_amino;          // parent code -> slots of (N7|C5|N1, H61|H41|H21, H62|H42|H22)

// ---------------------------------------------------------------- CIF tokens ----
struct Table {
    std::vector<std::string> cols;
    std::vector<sv> vals;              // row major
    size_t width() const { return cols.size(); }
    size_t rows() const { return cols.empty() ? 0 : vals.size() / cols.size(); }
    int col(const char *name) const {
        for (size_t k = 0; k < cols.size(); ++k)
            if (cols[k] == name) return (int)k;
        return -1;
    }
};

struct Tokens {
    std::string name;
    bool have_name = false;
    std::map<std::string, Table> tables;
};

struct WsTable {
    bool t[256];
    constexpr WsTable() : t() { t[(unsigned char)' '] = t[(unsigned char)'\t'] = t[(unsigned char)'\r'] = t[(unsigned char)'\n'] = true; }
};
constexpr WsTable g_ws;
inline bool is_ws(char c) { return g_ws.t[(unsigned char)c]; }

// general mmCIF tokenizer: data_ / loop_ / _names / values (quoted, ;-text fields, # comments), streamed straight into
// the tables (no intermediate token list: nearly all tokens of a file are atom_site values)
void tokenize(sv text, Tokens &out) {
    enum State { NONE, LOOP_NAMES, LOOP_VALUES, KV_VALUE };
    State state = NONE;
    Table *cur = nullptr;
    auto close_loop = [&]() {
        if (state == LOOP_VALUES && cur && cur->width()) cur->vals.resize(cur->vals.size() / cur->width() * cur->width());
        if (state == KV_VALUE && cur) cur->vals.push_back(sv("?"));
    };
    auto on_value = [&](sv v) {
        if ((state == LOOP_NAMES || state == LOOP_VALUES) && cur) { state = LOOP_VALUES; cur->vals.push_back(v); }
        else if (state == KV_VALUE && cur) { cur->vals.push_back(v); state = NONE; }
    };
    auto table_of = [&](sv nm, sv &attr) -> Table * {
        const size_t d = nm.find('.');
        attr = d == sv::npos ? nm : nm.substr(d + 1);
        return &out.tables[std::string(nm.substr(1, d == sv::npos ? sv::npos : d - 1))];
    };
    size_t i = 0, n = text.size();
    bool line_start = true;
    while (i < n) {
        const char c = text[i];
        if (c == '\n') { line_start = true; ++i; continue; }
        if (line_start && c == ';') {                       // text field up to a line that starts with ';'
            size_t j = i + 1, end = n;
            while (j < n) {
                const size_t nl = text.find('\n', j);
                if (nl == sv::npos) { j = n; break; }
                if (nl + 1 < n && text[nl + 1] == ';') { end = nl; j = nl + 2; break; }
                j = nl + 1;
            }
            sv v = text.substr(i + 1, (end == n ? n : end) - (i + 1));
            while (!v.empty() && is_ws(v.front())) v.remove_prefix(1);
            while (!v.empty() && is_ws(v.back())) v.remove_suffix(1);
            on_value(v);
            i = j;
            line_start = false;
            continue;
        }
        line_start = false;
        if (c == ' ' || c == '\t' || c == '\r') { ++i; continue; }
        if (c == '#') {
            while (i < n && text[i] != '\n') ++i;
            continue;
        }
        if (c == '\'' || c == '"') {
            size_t j = i + 1;
            while (j < n && text[j] != '\n' && !(text[j] == c && (j + 1 == n || is_ws(text[j + 1])))) ++j;
            on_value(text.substr(i + 1, j - i - 1));
            i = (j < n && text[j] == c) ? j + 1 : j;
            continue;
        }
        size_t j = i + 1;
        const char *p = text.data();
        while (j < n && !g_ws.t[(unsigned char)p[j]]) ++j;
        const sv t(p + i, j - i);
        i = j;
        if (c != '_' && c != 'l' && c != 'L' && c != 'd' && c != 'D') {      // the common case: a plain value
            on_value(t);
        } else if (t[0] == '_') {
            sv attr;
            if (state == LOOP_NAMES) {
                Table *tb = table_of(t, attr);
                if (!cur) {
                    cur = tb;
                    cur->cols.clear();
                    cur->vals.clear();
                    // atom_site holds nearly every token of a file (about one per 3.5 bytes): no regrowth
                    if (t.size() > 11 && t.compare(0, 11, "_atom_site.") == 0) cur->vals.reserve((text.size() - i) / 3 + 64);
                }
                cur->cols.emplace_back(attr);
            } else {
                close_loop();
                cur = table_of(t, attr);
                cur->cols.emplace_back(attr);
                state = KV_VALUE;
            }
        } else if (t.size() == 5 && strncasecmp(t.data(), "loop_", 5) == 0) {
            close_loop();
            state = LOOP_NAMES;
            cur = nullptr;
        } else if (t.size() >= 5 && (t[0] == 'd' || t[0] == 'D') && strncasecmp(t.data(), "data_", 5) == 0) {
            close_loop();
            state = NONE;
            cur = nullptr;
            if (!out.have_name) { out.name = std::string(t.substr(5)); out.have_name = true; }
        } else {
            on_value(t);
        }
    }
    close_loop();
}

bool parse_double(sv s, double &v) {
    const auto r = std::from_chars(s.data(), s.data() + s.size(), v);
    if (r.ec == std::errc() && r.ptr == s.data() + s.size()) return true;
    std::string tmp(s);                                   // "+1.5", "1.5(3)" and friends
    char *end = nullptr;
    v = strtod(tmp.c_str(), &end);
    return end != tmp.c_str();
}

bool parse_int(sv s, long long &v) {
    const auto r = std::from_chars(s.data(), s.data() + s.size(), v);
    return r.ec == std::errc() && r.ptr == s.data() + s.size();
}

// ---------------------------------------------------------------- one structure ----
struct Op {
    std::string id, name;
    double t[4][4];
};

struct NtOut {
    std::string model, chain, seq, sym, alt, ins;         // alt / ins: has_* say whether they are None
    bool has_alt = false, has_ins = false, has_index = false;
    long long number = 0, index = 0;
    double base[FR3D_BASE_SLOTS][3] = {}, bb[FR3D_BB_SLOTS][3] = {};
    uint32_t bmask = 0, bbmask = 0;
    int meta0 = -1, meta1 = 0, meta7 = 0;
    int prev = -1;                                        // nt of the structure whose O3' is this nt's previous O3'
    std::string uid;
};

struct StructOut {
    std::string pdb;
    std::vector<NtOut> nts;
    std::string error;
};

Op identity_op() {                                        // reader.py:146-160 (shift by (1, 1, 1))
    Op o;
    o.id = "I";
    o.name = "1_555";
    memset(o.t, 0, sizeof(o.t));
    for (int r = 0; r < 4; ++r) o.t[r][r] = 1.0;
    for (int r = 0; r < 3; ++r) o.t[r][3] = 1.0;
    return o;
}

std::string sym_name(const Op &o) {                       // reader.py:633-643
    if (o.name.empty() || o.name == "?") return "ASM_" + o.id;
    return o.name;
}

std::string strip_parens(sv s) {
    std::string o;
    for (char c : s)
        if (c != '(' && c != ')') o.push_back(c);
    return o;
}

std::vector<sv> split(sv s, char sep) {
    std::vector<sv> out;
    size_t a = 0;
    while (true) {
        const size_t b = s.find(sep, a);
        out.push_back(s.substr(a, b == sv::npos ? sv::npos : b - a));
        if (b == sv::npos) break;
        a = b + 1;
    }
    return out;
}

struct GenAtom {
    uint32_t row;       // atom_site row
    uint32_t sym;       // index into sym names
    double x, y, z;
};

void fill_atoms(NtOut &nt, const SeqInfo &si, const std::vector<std::pair<sv, const GenAtom *>> &atoms) {
    // packing._fill_atoms: duplicate names are averaged incrementally like AtomProxy (base.py:150-166)
    double b[FR3D_BASE_SLOTS][4] = {}, k[FR3D_BB_SLOTS][4] = {};
    bool hb[FR3D_BASE_SLOTS] = {}, hk[FR3D_BB_SLOTS] = {};
    bool saw_h = false;
    for (const auto &a : atoms) {
        const sv name = a.first;
        if (name.find('H') != sv::npos) saw_h = true;
        uint64_t key;
        if (!pack_name(name, key)) continue;
        double(*tgt)[4];
        bool *has;
        int s = si.base.find(key);
        if (s >= 0) { tgt = b; has = hb; }
        else {
            s = si.bb.find(key);
            if (s < 0) continue;
            tgt = k; has = hk;
        }
        const GenAtom *g = a.second;
        if (!has[s]) { tgt[s][0] = g->x; tgt[s][1] = g->y; tgt[s][2] = g->z; tgt[s][3] = 1; has[s] = true; }
        else {
            const double c = tgt[s][3];
            tgt[s][0] = (tgt[s][0] * c + g->x) / (c + 1);
            tgt[s][1] = (tgt[s][1] * c + g->y) / (c + 1);
            tgt[s][2] = (tgt[s][2] * c + g->z) / (c + 1);
            tgt[s][3] = c + 1;
        }
    }
    uint32_t bm = 0, km = 0;
    for (int s = 0; s < FR3D_BASE_SLOTS; ++s)
        if (hb[s]) { bm |= 1u << s; nt.base[s][0] = b[s][0]; nt.base[s][1] = b[s][1]; nt.base[s][2] = b[s][2]; }
    for (int s = 0; s < FR3D_BB_SLOTS; ++s)
        if (hk[s]) { km |= 1u << s; nt.bb[s][0] = k[s][0]; nt.bb[s][1] = k[s][1]; nt.bb[s][2] = k[s][2]; }
    int flags = si.flags, meta7 = 0;
    if ((flags & FR3D_FLAG_MODIFIED) && !saw_h) { flags |= FR3D_FLAG_DUP; meta7 = (int)si.dup; }
    nt.bmask = bm | (((uint32_t)(si.pcode + 1) & 15u) << 16) | (((uint32_t)flags & 15u) << 20);
    nt.bbmask = km;
    nt.meta0 = si.pcode; nt.meta1 = flags; nt.meta7 = meta7;
    if (!(si.flags & FR3D_FLAG_MODIFIED)) {                 // packing._fix_amino_hydrogens (components.py:405-445)
        auto am = g_amino.find(si.pcode);
        if (am != g_amino.end()) {
            const int sh = am->second.heavy, s1 = am->second.h1, s2 = am->second.h2;
            if ((bm >> sh & 1) && (bm >> s1 & 1) && (bm >> s2 & 1)) {
                double d1 = 0, d2 = 0;
                // (a0-p0)**2 + (a1-p1)**2 + (a2-p2)**2, left to right like the Python expression
                d1 = (nt.base[sh][0] - nt.base[s1][0]) * (nt.base[sh][0] - nt.base[s1][0]);
                d1 += (nt.base[sh][1] - nt.base[s1][1]) * (nt.base[sh][1] - nt.base[s1][1]);
                d1 += (nt.base[sh][2] - nt.base[s1][2]) * (nt.base[sh][2] - nt.base[s1][2]);
                d2 = (nt.base[sh][0] - nt.base[s2][0]) * (nt.base[sh][0] - nt.base[s2][0]);
                d2 += (nt.base[sh][1] - nt.base[s2][1]) * (nt.base[sh][1] - nt.base[s2][1]);
                d2 += (nt.base[sh][2] - nt.base[s2][2]) * (nt.base[sh][2] - nt.base[s2][2]);
                if (d1 < d2)
                    for (int c = 0; c < 3; ++c) std::swap(nt.base[s1][c], nt.base[s2][c]);
            }
        }
    }
}

std::string unit_id(const std::string &pdb, const NtOut &nt) {          // packing.encode_unit_id (fr3d/unit_ids.py:31-64)
    std::string f[9] = {pdb, nt.model, nt.chain, nt.seq, std::to_string(nt.number), "", nt.has_alt ? nt.alt : "",
                        nt.has_ins ? nt.ins : "", nt.sym};
    int n = 9;
    if (f[8] == "1_555") n = 8;
    while (n > 0 && f[n - 1].empty()) --n;
    std::string o;
    for (int k = 0; k < n; ++k) {
        if (k) o.push_back('|');
        o += f[k];
    }
    return o;
}

#ifdef FR3D_INGEST_TIMING
#define TICK(label) do { auto now_ = std::chrono::steady_clock::now(); fprintf(stderr, "  %-22s %.3f ms\n", label, std::chrono::duration<double, std::milli>(now_ - tick_).count()); tick_ = now_; } while (0)
#else
#define TICK(label)
#endif

void parse_structure(sv text, StructOut &out) {
#ifdef FR3D_INGEST_TIMING
    auto tick_ = std::chrono::steady_clock::now();
#endif
    Tokens T;
    tokenize(text, T);
    TICK("tokenize");
    out.pdb = T.name;
    auto at = T.tables.find("atom_site");
    if (at == T.tables.end() || at->second.rows() == 0) return;
    const Table &A = at->second;
    const size_t W = A.width(), R = A.rows();
    auto need = [&](const char *nm) { return A.col(nm); };
    const int c_asym = need("label_asym_id"), c_chain = need("auth_asym_id"), c_x = need("Cartn_x"), c_y = need("Cartn_y"),
              c_z = need("Cartn_z"), c_num = need("auth_seq_id");
    int c_comp = need("label_comp_id"), c_name = need("label_atom_id");
    if (c_comp < 0) c_comp = need("auth_comp_id");
    if (c_name < 0) c_name = need("auth_atom_id");
    const int c_model = need("pdbx_PDB_model_num"), c_ins = need("pdbx_PDB_ins_code"), c_alt = need("label_alt_id"),
              c_seq = need("label_seq_id");
    if (c_asym < 0 || c_chain < 0 || c_x < 0 || c_y < 0 || c_z < 0 || c_num < 0 || c_comp < 0 || c_name < 0) {
        out.error = "atom_site lacks a required column";
        return;
    }
    auto V = [&](size_t r, int c) { return A.vals[r * W + c]; };
    // chem_comp types
    std::unordered_map<std::string, bool> is_na;
    auto cc = T.tables.find("chem_comp");
    if (cc != T.tables.end()) {
        const Table &C = cc->second;
        const int ci = C.col("id"), ct = C.col("type");
        for (size_t r = 0; r < C.rows() && ci >= 0; ++r) {
            const sv ty = ct >= 0 ? C.vals[r * C.width() + ct] : sv();
            is_na[std::string(C.vals[r * C.width() + ci])] = (ty == "RNA linking" || ty == "DNA linking");
        }
    } else {
        for (const char *b : {"A", "C", "G", "U", "DA", "DC", "DG", "DT"}) is_na[b] = true;
    }
    // operators (reader.py:114-144)
    std::map<std::string, Op> ops;
    auto ol = T.tables.find("pdbx_struct_oper_list");
    if (ol != T.tables.end()) {
        const Table &O = ol->second;
        const int ci = O.col("id"), cn = O.col("name");
        for (size_t r = 0; r < O.rows() && ci >= 0; ++r) {
            Op o;
            o.id = std::string(O.vals[r * O.width() + ci]);
            o.name = cn >= 0 ? std::string(O.vals[r * O.width() + cn]) : std::string();
            memset(o.t, 0, sizeof(o.t));
            o.t[3][3] = 1.0;
            char nm[32];
            for (int a = 0; a < 3; ++a) {
                snprintf(nm, sizeof(nm), "vector[%d]", a + 1);
                int c = O.col(nm);
                if (c < 0 || !parse_double(O.vals[r * O.width() + c], o.t[a][3])) { out.error = "bad operator"; return; }
                for (int b = 0; b < 3; ++b) {
                    snprintf(nm, sizeof(nm), "matrix[%d][%d]", a + 1, b + 1);
                    c = O.col(nm);
                    if (c < 0 || !parse_double(O.vals[r * O.width() + c], o.t[a][b])) { out.error = "bad operator"; return; }
                }
            }
            ops[o.id] = o;
        }
    }
    ops["I"] = identity_op();
    // assemblies (reader.py:205-237, default path)
    std::vector<std::pair<std::string, std::vector<const Op *>>> assemblies;      // insertion order like the dict
    auto find_asm = [&](const std::string &a) -> std::vector<const Op *> * {
        for (auto &p : assemblies)
            if (p.first == a) return &p.second;
        return nullptr;
    };
    auto ag = T.tables.find("pdbx_struct_assembly_gen");
    if (ag != T.tables.end()) {
        const Table &G = ag->second;
        const int ce = G.col("oper_expression"), ca = G.col("asym_id_list");
        const Op *op = nullptr;
        for (size_t r = 0; r < G.rows() && ce >= 0 && ca >= 0; ++r) {
            const auto opers = split(G.vals[r * G.width() + ce], ',');
            for (sv asym_id : split(G.vals[r * G.width() + ca], ',')) {
                std::string a(asym_id);
                if (!find_asm(a)) assemblies.push_back({a, {}});
                auto *lst = find_asm(a);
                auto has = [&](const Op *o) {
                    for (const Op *x : *lst)
                        if (x->id == o->id) return true;
                    return false;
                };
                for (sv oper : opers) {
                    const bool hy = oper.find('-') != sv::npos, hx = oper.find('X') != sv::npos, hp = oper.find('P') != sv::npos;
                    if (hy && !hx && !hp) {
                        const std::string s = strip_parens(oper);
                        const auto ab = split(s, '-');
                        long long lo = 0, hi = -1;
                        if (ab.size() != 2 || !parse_int(ab[0], lo) || !parse_int(ab[1], hi)) { out.error = "bad operator range"; return; }
                        for (long long q = lo; q <= hi; ++q) {
                            auto it = ops.find(std::to_string(q));
                            if (it == ops.end()) { out.error = "unknown operator"; return; }
                            op = &it->second;
                            if (!has(op)) lst->push_back(op);
                        }
                        continue;
                    } else if (hx || hp) {
                        // nothing new is applied; the reference re-tests the operator of the previous iteration
                    } else {
                        auto it = ops.find(strip_parens(oper));
                        if (it == ops.end()) { out.error = "unknown operator"; return; }
                        op = &it->second;
                    }
                    if (!op) { out.error = "operator expression before any usable operator"; return; }
                    if (!has(op)) lst->push_back(op);
                }
            }
        }
    }
    // generate atoms: operator index outermost, file order inside (nucleotide residues only)
    std::vector<uint32_t> base_rows;
    base_rows.reserve(R);
    {
        std::string key;
        sv last;
        bool last_na = false;
        for (size_t r = 0; r < R; ++r) {
            const sv comp = V(r, c_comp);
            if (comp != last) {
                key.assign(comp);
                auto it = is_na.find(key);
                last_na = it != is_na.end() && it->second;
                last = comp;
            }
            if (last_na) base_rows.push_back((uint32_t)r);
        }
    }
    if (base_rows.empty()) return;
    TICK("tables, operators, rows");
    const Op ident = identity_op();
    std::vector<std::string> sym_names;
    auto sym_index = [&](const Op *o) {
        const std::string nm = sym_name(*o);
        for (size_t k = 0; k < sym_names.size(); ++k)
            if (sym_names[k] == nm) return (uint32_t)k;
        sym_names.push_back(nm);
        return (uint32_t)(sym_names.size() - 1);
    };
    size_t max_ops = 1;
    if (!assemblies.empty()) {
        max_ops = 0;
        for (auto &p : assemblies) max_ops = std::max(max_ops, p.second.size());
    }
    std::vector<double> X(R), Y(R), Z(R);
    for (uint32_t r : base_rows)
        if (!parse_double(V(r, c_x), X[r]) || !parse_double(V(r, c_y), Y[r]) || !parse_double(V(r, c_z), Z[r])) {
            out.error = "bad coordinate";
            return;
        }
    // residue numbers (int(auth_seq_id), digits only as the fallback, reader.py:590, 607-608)
    std::vector<long long> num(R, 0);
    for (uint32_t r : base_rows) {
        const sv s = V(r, c_num);
        if (!parse_int(s, num[r])) {
            std::string d;
            for (char ch : s)
                if (ch >= '0' && ch <= '9') d.push_back(ch);
            if (d.empty() || !parse_int(d, num[r])) { out.error = "bad auth_seq_id"; return; }
        }
    }
    const sv one("1"), qm("?"), dot(".");
    auto model_of = [&](uint32_t r) { return c_model >= 0 ? V(r, c_model) : one; };
    auto ins_of = [&](uint32_t r) { sv v = c_ins >= 0 ? V(r, c_ins) : qm; return v == qm ? sv() : v; };
    TICK("numbers");
    // Runs: maximal stretches of consecutive nucleotide rows with one residue key (and one label_asym_id, which picks
    // the operators).  Atoms of a run stay together and in file order under the reference's STABLE sort by residue key,
    // so sorting (operator index, run) pairs is the same as sorting the atoms - a few hundred keys instead of thousands.
    struct Run { uint32_t first, count; };            // indices into base_rows
    std::vector<Run> runs;
    for (uint32_t q = 0; q < base_rows.size(); ++q) {
        const uint32_t r = base_rows[q];
        bool same = false;
        if (!runs.empty()) {
            const uint32_t p = base_rows[q - 1];
            same = num[p] == num[r] && V(p, c_comp) == V(r, c_comp) && V(p, c_chain) == V(r, c_chain) &&
                   V(p, c_asym) == V(r, c_asym) && model_of(p) == model_of(r) && ins_of(p) == ins_of(r);
        }
        if (same) ++runs.back().count;
        else runs.push_back({q, 1});
    }
    struct GenRun { uint32_t run; const Op *op; uint32_t sym; };
    std::vector<GenRun> gen;                           // generation order: operator index outermost, file order inside
    gen.reserve(runs.size() * max_ops);
    for (size_t k = 0; k < max_ops; ++k) {
        for (uint32_t q = 0; q < runs.size(); ++q) {
            const uint32_t r = base_rows[runs[q].first];
            const Op *cur = nullptr;
            if (assemblies.empty()) cur = (k == 0) ? &ident : nullptr;
            else {
                auto *lst = find_asm(std::string(V(r, c_asym)));
                if (!lst || lst->empty()) cur = (k == 0) ? &ident : nullptr;                // reader.py:653-662
                else if (k < lst->size()) cur = (*lst)[k];
            }
            if (cur) gen.push_back({q, cur, sym_index(cur)});
        }
    }
    // stable sort by the residue key (strings compared like Python: by code point = by UTF-8 byte)
    std::vector<uint32_t> order(gen.size());
    for (uint32_t q = 0; q < order.size(); ++q) order[q] = q;
    auto cmp_key = [&](uint32_t a, uint32_t b) -> int {
        const uint32_t ra = base_rows[runs[gen[a].run].first], rb = base_rows[runs[gen[b].run].first];
        int c = model_of(ra).compare(model_of(rb));
        if (c) return c;
        c = V(ra, c_chain).compare(V(rb, c_chain));
        if (c) return c;
        c = V(ra, c_comp).compare(V(rb, c_comp));
        if (c) return c;
        if (num[ra] != num[rb]) return num[ra] < num[rb] ? -1 : 1;
        c = ins_of(ra).compare(ins_of(rb));
        if (c) return c;
        return sym_names[gen[a].sym].compare(sym_names[gen[b].sym]);
    };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cmp_key(a, b) < 0; });
    TICK("runs + sort");
    // groups -> conformers -> nucleotides
    std::string key;
    std::vector<GenAtom> atoms_buf;
    size_t a0 = 0;
    while (a0 < order.size()) {
        size_t a1 = a0 + 1;
        while (a1 < order.size() && cmp_key(order[a0], order[a1]) == 0) ++a1;
        // the atoms of the group, transformed: transform . (x, y, z, 1)  (reader.py:624-629)
        atoms_buf.clear();
        for (size_t q = a0; q < a1; ++q) {
            const GenRun &gr = gen[order[q]];
            const Run &run = runs[gr.run];
            for (uint32_t w = 0; w < run.count; ++w) {
                const uint32_t r = base_rows[run.first + w];
                GenAtom g;
                g.row = r;
                g.sym = gr.sym;
                g.x = gr.op->t[0][0] * X[r] + gr.op->t[0][1] * Y[r] + gr.op->t[0][2] * Z[r] + gr.op->t[0][3];
                g.y = gr.op->t[1][0] * X[r] + gr.op->t[1][1] * Y[r] + gr.op->t[1][2] * Z[r] + gr.op->t[1][3];
                g.z = gr.op->t[2][0] * X[r] + gr.op->t[2][1] * Y[r] + gr.op->t[2][2] * Z[r] + gr.op->t[2][3];
                atoms_buf.push_back(g);
            }
        }
        // alternate ids in order of first appearance
        std::vector<std::pair<sv, std::vector<uint32_t>>> groups;      // alt value ("" = None) -> members (atoms_buf indices)
        std::vector<bool> is_none;
        for (uint32_t q = 0; q < atoms_buf.size(); ++q) {
            sv v = c_alt >= 0 ? V(atoms_buf[q].row, c_alt) : dot;
            const bool none = (v == dot);
            if (none) v = sv();
            size_t g = 0;
            for (; g < groups.size(); ++g)
                if (groups[g].first == v && is_none[g] == none) break;
            if (g == groups.size()) { groups.push_back({v, {}}); is_none.push_back(none); }
            groups[g].second.push_back(q);
        }
        if (groups.size() > 1) {                                           // reader.py:428-438
            std::vector<uint32_t> common;
            for (size_t g = 0; g < groups.size(); ++g)
                if (is_none[g]) { common = groups[g].second; groups.erase(groups.begin() + g); is_none.erase(is_none.begin() + g); break; }
            for (auto &g : groups) g.second.insert(g.second.end(), common.begin(), common.end());
            std::stable_sort(groups.begin(), groups.end(), [](const auto &x, const auto &y) { return x.first < y.first; });
            is_none.assign(groups.size(), false);
        }
        const uint32_t r0 = atoms_buf[0].row;
        for (size_t g = 0; g < groups.size(); ++g) {
            out.nts.emplace_back();
            NtOut &nt = out.nts.back();
            nt.model = std::string(model_of(r0));
            nt.chain = std::string(V(r0, c_chain));
            nt.seq = std::string(V(r0, c_comp));
            nt.number = num[r0];
            const sv sq = c_seq >= 0 ? V(r0, c_seq) : dot;
            nt.has_index = !(sq == dot || sq.empty() || sq == qm) && parse_int(sq, nt.index);
            nt.sym = sym_names[gen[order[a0]].sym];
            nt.has_alt = !is_none[g];
            nt.alt = std::string(groups[g].first);
            const sv iv = ins_of(r0);
            nt.has_ins = !iv.empty();
            nt.ins = std::string(iv);
            key.assign(nt.seq);
            auto si = g_seq.find(key);
            if (si != g_seq.end()) {
                std::vector<std::pair<sv, const GenAtom *>> atoms;
                atoms.reserve(groups[g].second.size());
                for (uint32_t q : groups[g].second) atoms.push_back({V(atoms_buf[q].row, c_name), &atoms_buf[q]});
                fill_atoms(nt, si->second, atoms);
            }
            nt.uid = unit_id(out.pdb, nt);
        }
        a0 = a1;
    }
    TICK("residues + slots");
    // previous O3' (map_unit_id_to_previous_O3, NA:480-525; packing._prev_o3_of_structure): the nts of the structure in
    // (model, symmetry, chain, index, unit id) order, successive indices of one chain are linked
    std::vector<int> ord(out.nts.size());
    for (size_t q = 0; q < ord.size(); ++q) ord[q] = (int)q;
    std::sort(ord.begin(), ord.end(), [&](int a, int b) {
        const NtOut &x = out.nts[a], &y = out.nts[b];
        int c = x.model.compare(y.model);
        if (c) return c < 0;
        c = x.sym.compare(y.sym);
        if (c) return c < 0;
        c = x.chain.compare(y.chain);
        if (c) return c < 0;
        const long long ix = x.has_index ? x.index : 0, iy = y.has_index ? y.index : 0;
        if (ix != iy) return ix < iy;
        c = x.uid.compare(y.uid);
        if (c) return c < 0;
        return a < b;
    });
    for (size_t q = 1; q < ord.size(); ++q) {
        const NtOut &x = out.nts[ord[q - 1]];
        NtOut &y = out.nts[ord[q]];
        const long long ix = x.has_index ? x.index : 0, iy = y.has_index ? y.index : 0;
        if (x.model == y.model && x.sym == y.sym && x.chain == y.chain && iy == ix + 1 && x.meta0 >= 0 && (x.bbmask >> 5 & 1u))
            y.prev = ord[q - 1];
    }
}

}  // namespace

struct fr3d_ingest {
    std::vector<StructOut> structs;
    int64_t n_nt = 0;
    std::string columns[7];            // model chain sequence symmetry alt_id insertion_code pdb, '\0' separated ("\1" = None)
};

extern "C" {

const char *fr3d_ingest_last_error(void) { return g_ingest_err; }

int fr3d_ingest_tables(const char *blob, size_t len) {
    if (!blob) return ifail("blob is NULL");
    g_seq.clear();
    g_amino.clear();
    sv text(blob, len);
    SeqInfo *cur = nullptr;
    size_t a = 0;
    while (a < text.size()) {
        size_t b = text.find('\n', a);
        if (b == sv::npos) b = text.size();
        sv line = text.substr(a, b - a);
        a = b + 1;
        if (line.empty()) continue;
        std::vector<sv> f = split(line, '\t');
        if (f[0] == "S" && f.size() == 5) {
            long long p, fl, d;
            if (!parse_int(f[2], p) || !parse_int(f[3], fl) || !parse_int(f[4], d)) return ifail("bad S line");
            cur = &g_seq[std::string(f[1])];
            cur->pcode = (int)p; cur->flags = (int)fl; cur->dup = (uint32_t)d;
        } else if ((f[0] == "B" || f[0] == "K") && f.size() == 3 && cur) {
            long long s;
            if (!parse_int(f[2], s)) return ifail("bad map line");
            uint64_t key;
            if (!pack_name(f[1], key)) return ifail("atom name longer than 8 bytes in the slot tables");
            (f[0] == "B" ? cur->base : cur->bb).set(key, (int)s);
        } else if (f[0] == "A" && f.size() == 5) {
            long long p, h, h1, h2;
            if (!parse_int(f[1], p) || !parse_int(f[2], h) || !parse_int(f[3], h1) || !parse_int(f[4], h2)) return ifail("bad A line");
            g_amino[(int)p] = {(int)h, (int)h1, (int)h2};
        } else {
            return ifail("bad table line");
        }
    }
    return FR3D_OK;
}

int fr3d_cif_parse(const char *const *texts, const int64_t *lengths, int32_t n_texts, int32_t n_threads, fr3d_ingest **out) {
    if (!texts || !lengths || !out || n_texts < 0) return ifail("bad argument");
    if (g_seq.empty()) return ifail("fr3d_ingest_tables has not been called");
    std::unique_ptr<fr3d_ingest> H(new fr3d_ingest);
    H->structs.resize(n_texts);
    std::atomic<int> next(0);
    auto work = [&]() {
        while (true) {
            const int k = next.fetch_add(1);
            if (k >= n_texts) break;
            try {
                parse_structure(sv(texts[k], (size_t)lengths[k]), H->structs[k]);
            } catch (const std::exception &ex) {
                H->structs[k].error = ex.what();
            }
        }
    };
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, n_texts));
    if (nt == 1) work();
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(work);
        for (auto &t : th) t.join();
    }
    for (int k = 0; k < n_texts; ++k)
        if (!H->structs[k].error.empty()) {
            snprintf(g_ingest_err, sizeof(g_ingest_err), "structure %d: %s", k, H->structs[k].error.c_str());
            return FR3D_E_ARG;
        }
    int64_t n = 0;
    for (auto &s : H->structs) n += (int64_t)s.nts.size();
    H->n_nt = n;
    auto put = [](std::string &col, const std::string &v, bool present) {
        if (present) col += v; else col.push_back('\1');
        col.push_back('\0');
    };
    for (auto &s : H->structs) {
        put(H->columns[6], s.pdb, true);
        for (auto &nt : s.nts) {
            put(H->columns[0], nt.model, true);
            put(H->columns[1], nt.chain, true);
            put(H->columns[2], nt.seq, true);
            put(H->columns[3], nt.sym, true);
            put(H->columns[4], nt.alt, nt.has_alt);
            put(H->columns[5], nt.ins, nt.has_ins);
        }
    }
    *out = H.release();
    return FR3D_OK;
}

// sizes[0] = nts, [1] = structures, [2..8] = bytes of the seven string columns
int fr3d_ingest_sizes(const fr3d_ingest *h, int64_t *sizes) {
    if (!h || !sizes) return ifail("bad argument");
    sizes[0] = h->n_nt;
    sizes[1] = (int64_t)h->structs.size();
    for (int k = 0; k < 7; ++k) sizes[2 + k] = (int64_t)h->columns[k].size();
    return FR3D_OK;
}

// Fills the record planes (host pointers, e.g. pinned memory; base_xyz [n][16][3], bb_xyz [n][10][3], masks [n], meta [n][8]),
// struct_off [S + 1], number / index [n] (index INT64_MIN = None) and the string columns.  The identity part of meta
// (group, chain rank, residue key, alt rank) and the previous-O3' slot follow packing._finish_identity.
int fr3d_ingest_fill(const fr3d_ingest *h, double *base_xyz, double *bb_xyz, uint32_t *base_mask, uint32_t *bb_mask,
                     int32_t *meta, int64_t *struct_off, int64_t *number, int64_t *index, char *const *string_columns) {
    if (!h || !base_xyz || !bb_xyz || !base_mask || !bb_mask || !meta || !struct_off || !number || !index || !string_columns)
        return ifail("bad argument");
    const int64_t n = h->n_nt;
    for (int k = 0; k < 7; ++k)
        if (string_columns[k]) memcpy(string_columns[k], h->columns[k].data(), h->columns[k].size());
    std::vector<const NtOut *> nts;
    std::vector<int32_t> sof;
    nts.reserve(n);
    sof.reserve(n);
    int64_t off = 0;
    for (size_t s = 0; s < h->structs.size(); ++s) {
        struct_off[s] = off;
        for (auto &nt : h->structs[s].nts) { nts.push_back(&nt); sof.push_back((int32_t)s); }
        off += (int64_t)h->structs[s].nts.size();
    }
    struct_off[h->structs.size()] = off;
    // chain ranks over the whole batch
    std::map<std::string, int> chain_rank;
    for (auto *nt : nts) chain_rank[nt->chain] = 0;
    {
        int r = 0;
        for (auto &p : chain_rank) p.second = r++;
    }
    bool any_alt = false;
    for (auto *nt : nts) any_alt = any_alt || nt->has_alt;
    std::map<std::pair<int, std::string>, int> groups;
    std::unordered_map<std::string, int> resid, alts;
    int last_sof = -1, last_group = 0;
    const std::string *last_model = nullptr, *last_chain = nullptr;
    int last_chain_rank = 0;
    std::string key;
    for (int64_t k = 0; k < n; ++k) {
        const NtOut &nt = *nts[k];
        memcpy(base_xyz + k * FR3D_BASE_SLOTS * 3, nt.base, sizeof(nt.base));
        memcpy(bb_xyz + k * FR3D_BB_SLOTS * 3, nt.bb, sizeof(nt.bb));
        base_mask[k] = nt.meta0 >= 0 ? nt.bmask : 0;
        bb_mask[k] = nt.meta0 >= 0 ? nt.bbmask : 0;
        int32_t *m = meta + k * FR3D_META_INTS;
        m[0] = nt.meta0; m[1] = nt.meta1; m[7] = nt.meta7;
        if (!(last_sof == sof[k] && last_model && *last_model == nt.model)) {      // consecutive nts mostly share both
            auto gk = std::make_pair(sof[k], nt.model);
            auto gi = groups.find(gk);
            if (gi == groups.end()) gi = groups.emplace(gk, (int)groups.size()).first;
            last_group = gi->second; last_sof = sof[k]; last_model = &nt.model;
        }
        m[2] = last_group;
        if (!(last_chain && *last_chain == nt.chain)) { last_chain_rank = chain_rank[nt.chain]; last_chain = &nt.chain; }
        m[3] = last_chain_rank;
        m[4] = nt.has_index ? (int32_t)nt.index : 0;
        if (any_alt) {
            key = std::to_string(sof[k]) + "\x1f" + std::to_string(nt.number) + "\x1f" + nt.seq + "\x1f" + nt.chain + "\x1f" + nt.sym +
                  "\x1f" + (nt.has_ins ? nt.ins : std::string("\1"));
            auto ri = resid.find(key);
            if (ri == resid.end()) ri = resid.emplace(key, (int)resid.size()).first;
            m[5] = ri->second;
            key = nt.has_alt ? nt.alt : std::string("\1");
            auto ai = alts.find(key);
            if (ai == alts.end()) ai = alts.emplace(key, (int)alts.size()).first;
            m[6] = ai->second;
        } else {
            m[5] = (int32_t)k;
            m[6] = 0;
        }
        number[k] = nt.number;
        index[k] = nt.has_index ? nt.index : INT64_MIN;
    }
    // previous O3' (found per structure by the parsing threads)
    for (int64_t k = 0; k < n; ++k) {
        const NtOut &nt = *nts[k];
        if (nt.prev >= 0) {
            const NtOut &x = h->structs[sof[k]].nts[nt.prev];
            memcpy(bb_xyz + (k * FR3D_BB_SLOTS + 9) * 3, x.bb[5], 3 * sizeof(double));
            bb_mask[k] |= 1u << 9;
        }
    }
    return FR3D_OK;
}

void fr3d_ingest_free(fr3d_ingest *h) { delete h; }

}  // extern "C"
