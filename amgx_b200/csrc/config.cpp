// config.cpp -- see config.h.  Own JSON reader (the reference vendors rapidjson; we need only
// objects / strings / numbers / true / false / null / arrays-skipped).
#include "config.h"
#include <fstream>
#include <sstream>
#include <algorithm>
#include <cctype>
#include <cmath>

namespace amgxb {

// ---------------------------------------------------------------------------------------------
// Parameter registry: names, types and defaults of the reference's registerParameters()
// (src/core.cu:307-543).  Enum-typed parameters (algorithm, norm, view types, colouring types,
// block format) are kept as strings.  Everything registered there parses here; parameters that
// do not concern the solve-phase engine are accepted and ignored.
// ---------------------------------------------------------------------------------------------
static const ParamDesc g_registry[] = {
    {"determinism_flag", PType::INT, "0"},
    {"exception_handling", PType::INT, "0"},
    {"fine_level_consolidation", PType::INT, "0"},
    {"use_cuda_ipc_consolidation", PType::INT, "0"},
    {"amg_consolidation_flag", PType::INT, "0"},
    {"matrix_consolidation_lower_threshold", PType::INT, "0"},
    {"matrix_consolidation_upper_threshold", PType::INT, "1000"},
    {"device_mem_pool_size", PType::SIZE, "268435456"},
    {"device_consolidation_pool_size", PType::SIZE, "268435456"},
    {"device_mem_pool_max_alloc_size", PType::SIZE, "20971520"},
    {"device_alloc_scaling_factor", PType::SIZE, "10"},
    {"device_alloc_scaling_threshold", PType::SIZE, "16384"},
    {"device_mem_pool_size_limit", PType::SIZE, "0"},
    {"num_streams", PType::INT, "0"},
    {"serialize_threads", PType::INT, "0"},
    {"high_priority_stream", PType::INT, "0"},
    {"communicator", PType::STRING, "MPI"},
    {"separation_interior", PType::STRING, "INTERIOR"},
    {"separation_exterior", PType::STRING, "OWNED"},
    {"min_rows_latency_hiding", PType::INT, "-1"},
    {"exact_coarse_solve", PType::INT, "0"},
    {"matrix_halo_exchange", PType::INT, "0"},
    {"boundary_coloring", PType::STRING, "SYNC_COLORS"},
    {"halo_coloring", PType::STRING, "LAST"},
    {"use_sum_stopping_criteria", PType::INT, "0"},
    {"rhs_from_a", PType::INT, "0"},
    {"complex_conversion", PType::INT, "0"},
    {"matrix_writer", PType::STRING, "matrixmarket"},
    {"block_format", PType::STRING, "ROW_MAJOR"},
    {"block_convert", PType::INT, "0"},
    {"solver", PType::STRING, "AMG"},
    {"preconditioner", PType::STRING, "AMG"},
    {"coarse_solver", PType::STRING, "DENSE_LU_SOLVER"},
    {"smoother", PType::STRING, "BLOCK_JACOBI"},
    {"fine_smoother", PType::STRING, "BLOCK_JACOBI"},
    {"coarse_smoother", PType::STRING, "BLOCK_JACOBI"},
    {"gmres_n_restart", PType::INT, "20"},
    {"gmres_krylov_dim", PType::INT, "0"},
    {"subspace_dim_s", PType::INT, "8"},
    {"dense_lu_num_rows", PType::INT, "128"},
    {"dense_lu_max_rows", PType::INT, "0"},
    {"relaxation_factor", PType::DOUBLE, "0.9"},
    {"ilu_sparsity_level", PType::INT, "0"},
    {"symmetric_GS", PType::INT, "0"},
    {"jacobi_iters", PType::INT, "5"},
    {"GS_L1_variant", PType::INT, "0"},
    {"kpz_mu", PType::INT, "4"},
    {"kpz_order", PType::INT, "3"},
    {"chebyshev_polynomial_order", PType::INT, "5"},
    {"chebyshev_lambda_estimate_mode", PType::INT, "0"},
    {"cheby_max_lambda", PType::DOUBLE, "1.0"},
    {"cheby_min_lambda", PType::DOUBLE, "0.125"},
    {"kaczmarz_coloring_needed", PType::INT, "1"},
    {"cf_smoothing_mode", PType::INT, "0"},
    {"algorithm", PType::STRING, "CLASSICAL"},
    {"amg_host_levels_rows", PType::INT, "-1"},
    {"cycle", PType::STRING, "V"},
    {"max_levels", PType::INT, "100"},
    {"min_fine_rows", PType::INT, "1"},
    {"min_coarse_rows", PType::INT, "2"},
    {"max_coarse_iters", PType::INT, "100"},
    {"coarsen_threshold", PType::DOUBLE, "1.0"},
    {"presweeps", PType::INT, "1"},
    {"postsweeps", PType::INT, "1"},
    {"finest_sweeps", PType::INT, "-1"},
    {"coarsest_sweeps", PType::INT, "2"},
    {"cycle_iters", PType::INT, "2"},
    {"structure_reuse_levels", PType::INT, "0"},
    {"error_scaling", PType::INT, "0"},
    {"reuse_scale", PType::INT, "0"},
    {"scaling_smoother_steps", PType::INT, "2"},
    {"intensive_smoothing", PType::INT, "0"},
    {"coarseAgenerator", PType::STRING, "LOW_DEG"},
    {"coarseAgenerator_coarse", PType::STRING, "LOW_DEG"},
    {"interpolator", PType::STRING, "D1"},
    {"energymin_interpolator", PType::STRING, "EM"},
    {"energymin_selector", PType::STRING, "CR"},
    {"selector", PType::STRING, "PMIS"},
    {"aggressive_levels", PType::INT, "0"},
    {"aggressive_selector", PType::STRING, "DEFAULT"},
    {"aggressive_interpolator", PType::STRING, "MULTIPASS"},
    {"handshaking_phases", PType::INT, "1"},
    {"aggregation_edge_weight_component", PType::INT, "0"},
    {"max_matching_iterations", PType::INT, "15"},
    {"max_unassigned_percentage", PType::DOUBLE, "0.05"},
    {"weight_formula", PType::INT, "0"},
    {"aggregation_passes", PType::INT, "3"},
    {"filter_weights", PType::INT, "0"},
    {"filter_weights_alpha", PType::DOUBLE, "0.5"},
    {"full_ghost_level", PType::INT, "0"},
    {"notay_weights", PType::INT, "0"},
    {"ghost_offdiag_limit", PType::INT, "0"},
    {"merge_singletons", PType::INT, "1"},
    {"serial_matching", PType::INT, "0"},
    {"modified_handshake", PType::INT, "0"},
    {"aggregate_size", PType::INT, "2"},
    {"strength", PType::STRING, "AHAT"},
    {"strength_threshold", PType::DOUBLE, "0.25"},
    {"max_row_sum", PType::DOUBLE, "1.1"},
    {"interp_truncation_factor", PType::DOUBLE, "1.1"},
    {"interp_max_elements", PType::INT, "-1"},
    {"affinity_iterations", PType::INT, "4"},
    {"affinity_vectors", PType::INT, "4"},
    {"coloring_level", PType::INT, "1"},
    {"reorder_cols_by_color", PType::INT, "0"},
    {"insert_diag_while_reordering", PType::INT, "0"},
    {"matrix_coloring_scheme", PType::STRING, "MIN_MAX"},
    {"max_num_hash", PType::INT, "7"},
    {"num_colors", PType::INT, "10"},
    {"max_uncolored_percentage", PType::DOUBLE, "0.15"},
    {"initial_color", PType::INT, "0"},
    {"use_bsrxmv", PType::INT, "0"},
    {"fine_levels", PType::INT, "-1"},
    {"coloring_try_remove_last_colors", PType::INT, "0"},
    {"coloring_custom_arg", PType::STRING, ""},
    {"print_coloring_info", PType::INT, "0"},
    {"weakness_bound", PType::INT, "2147483647"},
    {"late_rejection", PType::INT, "0"},
    {"geometric_dim", PType::INT, "2"},
    {"spmm_gmem_size", PType::INT, "1024"},
    {"spmm_no_sort", PType::INT, "1"},
    {"spmm_verbose", PType::INT, "0"},
    {"spmm_max_attempts", PType::INT, "6"},
    {"use_opt_kernels", PType::INT, "0"},
    {"use_cusparse_spgemm", PType::INT, "0"},
    {"cusparse_spgemm_alg", PType::STRING, "CUSPARSE_SPGEMM_DEFAULT"},
    {"cusparse_spgemm_fraction", PType::DOUBLE, "0.5"},
    {"max_iters", PType::INT, "100"},
    {"monitor_residual", PType::INT, "0"},
    {"convergence", PType::STRING, "ABSOLUTE"},
    {"norm", PType::STRING, "L2"},
    {"use_scalar_norm", PType::INT, "0"},
    {"tolerance", PType::DOUBLE, "1e-12"},
    {"alt_rel_tolerance", PType::DOUBLE, "1e-12"},
    {"rel_div_tolerance", PType::DOUBLE, "-1"},
    {"verbosity_level", PType::INT, "3"},
    {"solver_verbose", PType::INT, "0"},
    {"print_config", PType::INT, "0"},
    {"print_solve_stats", PType::INT, "0"},
    {"print_grid_stats", PType::INT, "0"},
    {"print_vis_data", PType::INT, "0"},
    {"print_aggregation_info", PType::INT, "0"},
    {"obtain_timings", PType::INT, "0"},
    {"store_res_history", PType::INT, "0"},
    {"convergence_analysis", PType::INT, "0"},
    {"scaling", PType::STRING, "NONE"},
    // eigensolver parameters registered by the reference's eigen registry: accepted, ignored
    {"eig_solver", PType::STRING, "POWER_ITERATION"},
    {"eig_max_iters", PType::INT, "100"},
    {"eig_tolerance", PType::DOUBLE, "1e-4"},
    {"eig_shift", PType::DOUBLE, "0"},
    {"eig_damping_factor", PType::DOUBLE, "0.85"},
    {"eig_which", PType::STRING, "largest"},
    {"eig_eigenvector", PType::INT, "0"},
    {"eig_eigenvector_solver", PType::STRING, "default"},
    {"eig_subspace_size", PType::INT, "2"},
    {"eig_wanted_count", PType::INT, "1"},
};

const ParamDesc *Config::registry(size_t *count)
{
    *count = sizeof(g_registry) / sizeof(g_registry[0]);
    return g_registry;
}

const ParamDesc *Config::find_desc(const std::string &name)
{
    size_t n;
    const ParamDesc *r = registry(&n);
    for (size_t i = 0; i < n; i++)
        if (name == r[i].name) return &r[i];
    return nullptr;
}

// names that may carry a new scope (src/amg_config.cu: m_solver_list)
static bool is_solver_param(const std::string &n)
{
    return n == "solver" || n == "preconditioner" || n == "smoother" || n == "coarse_solver" ||
           n == "fine_smoother" || n == "coarse_smoother" || n == "eig_solver" || n == "eig_eigenvector_solver";
}

static std::string trim(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && isspace((unsigned char)s[a])) a++;
    while (b > a && isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

static bool valid_token(const std::string &s)   // allowed_symbol() in the reference
{
    if (s.empty()) return false;
    for (char c : s)
        if (!(isalnum((unsigned char)c) || c == '_' || c == '.' || c == '-' || c == '+' || c == '/' || c == '~'))
            return false;
    return true;
}

static ParamValue make_value(const ParamDesc *d, const std::string &text)
{
    ParamValue v;
    v.type = d->type;
    try {
        switch (d->type) {
        case PType::INT:
        case PType::SIZE: {
            size_t pos = 0;
            // the reference accepts "1e3"-style ints only through JSON doubles; legacy strings use stream extraction
            v.i = std::stoll(text, &pos);
            if (pos != text.size()) {
                double dd = std::stod(text, &pos);
                if (pos != text.size()) throw std::invalid_argument("x");
                v.i = (long long)dd;
            }
            break;
        }
        case PType::DOUBLE: {
            size_t pos = 0;
            v.d = std::stod(text, &pos);
            if (pos != text.size()) throw std::invalid_argument("x");
            break;
        }
        case PType::STRING: v.s = text; break;
        }
    } catch (...) {
        fatal(AMGX_RC_BAD_CONFIGURATION, std::string("cannot convert value '") + text + "' of parameter '" + d->name + "'");
    }
    return v;
}

void Config::import_named(const std::string &name, const std::string &text, bool is_string_token,
                          bool is_double_token, const std::string &cur_scope, const std::string &new_scope)
{
    (void)is_double_token;
    if (std::find(scopes_.begin(), scopes_.end(), new_scope) == scopes_.end())
        scopes_.push_back(new_scope);
    else if (new_scope != "default" && !allow_mod)
        fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (new scope already defined): " + new_scope);

    const ParamDesc *d = find_desc(name);
    if (!d) fatal(AMGX_RC_BAD_CONFIGURATION, "Variable '" + name + "' not registered");

    static const char *default_only[] = {"determinism_flag", "block_format", "separation_interior", "separation_exterior",
                                         "min_rows_latency_hiding", "fine_level_consolidation", "use_cuda_ipc_consolidation"};
    for (const char *n : default_only)
        if (name == n && cur_scope != "default")
            fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry. Parameter " + name + " can only be specified with default scope.");

    if (new_scope != "default" && !is_solver_param(name))
        fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry. New scope can only be associated with a solver. new_scope=" +
                                             new_scope + ", name=" + name + ".");

    if (is_string_token && d->type != PType::STRING)
        fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry. Type of the parameter \"" + name + "\" in the config is string");
    if (!is_string_token && d->type == PType::STRING && false) {}

    ParamValue v = make_value(d, text);
    v.new_scope = new_scope;
    params_[{cur_scope, name}] = v;
}

// ---------------------------------------------------------------------------------------------
// legacy "config_version=2, scope:name(new_scope)=value, ..." format
// ---------------------------------------------------------------------------------------------
void Config::set_one_legacy(const std::string &entry)
{
    if (std::count(entry.begin(), entry.end(), '=') != 1)
        fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (number of equal signs is not 1) : " + entry);
    size_t eq = entry.find('=');
    std::string value = trim(entry.substr(eq + 1));
    std::string name = entry.substr(0, eq);
    std::string new_scope = "default", cur_scope = "default";
    int nl = (int)std::count(name.begin(), name.end(), '('), nr = (int)std::count(name.begin(), name.end(), ')');
    if (nl != nr || nl > 1)
        fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (incorrect number of parentheses or unbalanced parantheses): " + entry);
    if (nl == 1) {
        size_t l = name.find('('), r = name.find(')');
        new_scope = trim(name.substr(l + 1, r - l - 1));
        name = name.substr(0, l);
        if (!valid_token(new_scope))
            fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (invalid symbol or empty string after trimming new_scope): " + entry);
        if (new_scope == "default")
            fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (new scope cannot be default scope): " + entry);
    }
    int nc = (int)std::count(name.begin(), name.end(), ':');
    if (nc > 1) fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (number of colons is > 1): " + entry);
    if (nc == 1) {
        size_t c = name.find(':');
        cur_scope = trim(name.substr(0, c));
        name = name.substr(c + 1);
        if (!valid_token(cur_scope))
            fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (invalid string or empty string after trimming current_scope): " + entry);
    }
    name = trim(name);
    if (!valid_token(name) || !valid_token(value))
        fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry (invalid string or empty string after stripping name or value): " + entry);
    const ParamDesc *d = find_desc(name);
    if (!d) fatal(AMGX_RC_BAD_CONFIGURATION, "Variable '" + name + "' not registered");
    import_named(name, value, /*is_string_token=*/d->type == PType::STRING, false, cur_scope, new_scope);
}

void Config::parse_legacy(std::string params)
{
    // split on ',' or ';'
    std::vector<std::string> entries;
    std::string cur;
    for (char c : params) {
        if (c == ',' || c == ';') { entries.push_back(cur); cur.clear(); }
        else cur += c;
    }
    entries.push_back(cur);
    int version = 1;
    size_t first = 0;
    // config_version must be the first entry if present (src/amg_config.cu:150-185)
    if (!entries.empty() && entries[0].size() > 2 && trim(entries[0]).size()) {
        std::string e = entries[0];
        size_t eq = e.find('=');
        if (eq != std::string::npos && trim(e.substr(0, eq)) == "config_version") {
            version = atoi(trim(e.substr(eq + 1)).c_str());
            if (version != 1 && version != 2)
                fatal(AMGX_RC_BAD_CONFIGURATION, "Error, config_version must be 1 or 2. Config string is " + e);
            first = 1;
        }
    }
    for (size_t k = first; k < entries.size(); k++) {
        std::string e = entries[k];
        if (e.size() <= 2 || trim(e).empty()) continue;
        if (version == 1) {
            // v1 -> v2 conversion (src/amg_config.cu:185-250): no scopes allowed, a few renames
            if (e.find(':') != std::string::npos || e.find('(') != std::string::npos)
                fatal(AMGX_RC_BAD_CONFIGURATION, "Error parsing parameter string: " + e +
                      " . Scopes only supported with config_version=2 and higher. Add \"config_version=2\" to the config string to use nested solvers");
            size_t eq = e.find('=');
            if (eq != std::string::npos) {
                std::string n = trim(e.substr(0, eq)), v = trim(e.substr(eq + 1));
                if (n == "smoother_weight") e = "relaxation_factor=" + v;
                else if (n == "min_block_rows") e = "min_coarse_rows=" + v;
                else if (v == "JACOBI" || v == "JACOBI_NO_CUSP") e = n + "=BLOCK_JACOBI";
            }
        }
        set_one_legacy(e);
    }
}

// ---------------------------------------------------------------------------------------------
// minimal JSON reader -> tree
// ---------------------------------------------------------------------------------------------
namespace {
struct JVal {
    enum Kind { NUL, BOOL, INT, DBL, STR, OBJ, ARR } kind = NUL;
    bool b = false;
    long long i = 0;
    double d = 0;
    std::string s, raw;
    std::vector<std::pair<std::string, JVal>> members;   // insertion order kept
    const JVal *get(const std::string &k) const
    {
        for (auto &m : members) if (m.first == k) return &m.second;
        return nullptr;
    }
};
struct JParser {
    const char *p, *end;
    bool ok = true;
    void ws() { while (p < end && isspace((unsigned char)*p)) p++; }
    bool str(std::string &out)
    {
        if (p >= end || *p != '"') return false;
        p++;
        out.clear();
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
                case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                case 'u': { if (end - p < 5) return false; out += '?'; p += 4; break; }
                default: out += *p;
                }
                p++;
            } else out += *p++;
        }
        if (p >= end) return false;
        p++;
        return true;
    }
    bool value(JVal &v)
    {
        ws();
        if (p >= end) return false;
        if (*p == '{') {
            v.kind = JVal::OBJ; p++; ws();
            if (p < end && *p == '}') { p++; return true; }
            while (true) {
                ws();
                std::string k;
                if (!str(k)) return false;
                ws();
                if (p >= end || *p != ':') return false;
                p++;
                JVal c;
                if (!value(c)) return false;
                v.members.emplace_back(k, std::move(c));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; return true; }
                return false;
            }
        }
        if (*p == '[') {
            v.kind = JVal::ARR; p++; ws();
            if (p < end && *p == ']') { p++; return true; }
            while (true) {
                JVal c;
                if (!value(c)) return false;
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; return true; }
                return false;
            }
        }
        if (*p == '"') { v.kind = JVal::STR; return str(v.s); }
        if (!strncmp(p, "true", 4) && end - p >= 4) { v.kind = JVal::BOOL; v.b = true; p += 4; return true; }
        if (!strncmp(p, "false", 5) && end - p >= 5) { v.kind = JVal::BOOL; v.b = false; p += 5; return true; }
        if (!strncmp(p, "null", 4) && end - p >= 4) { v.kind = JVal::NUL; p += 4; return true; }
        // number
        const char *s = p;
        if (p < end && (*p == '-' || *p == '+')) p++;
        bool isd = false, digits = false;
        while (p < end && (isdigit((unsigned char)*p) || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+')) {
            if (*p == '.' || *p == 'e' || *p == 'E') isd = true;
            if (isdigit((unsigned char)*p)) digits = true;
            p++;
        }
        if (!digits) return false;
        v.raw.assign(s, p);
        if (isd) { v.kind = JVal::DBL; v.d = strtod(v.raw.c_str(), nullptr); }
        else { v.kind = JVal::INT; v.i = strtoll(v.raw.c_str(), nullptr, 10); v.d = (double)v.i; }
        return true;
    }
};
}  // namespace

// import_json_object of the reference (src/amg_config.cu:545-610): nested objects are solvers
// living in their own scope ("scope" member, else "<current>_sub_<name>").
static void import_json(Config &cfg, const JVal &obj, bool outer,
                        void (*imp)(Config &, const std::string &, const JVal &, const std::string &, const std::string &))
{
    std::string cur = "default";
    if (const JVal *s = obj.get("scope")) if (s->kind == JVal::STR) cur = s->s;
    for (auto &m : obj.members) {
        const std::string &name = m.first;
        const JVal &v = m.second;
        if (name == "config_version" || name == "scope") continue;
        if ((name == "solver" || name == "eig_solver") && !outer && v.kind != JVal::OBJ) continue;
        if ((name == "solver" || name == "eig_solver") && !outer) continue;
        if (v.kind == JVal::OBJ) {
            std::string sub = cur + "_sub_" + name;
            if (const JVal *s = v.get("scope")) if (s->kind == JVal::STR) sub = s->s;
            const JVal *sv = v.get("solver");
            if (!sv || sv->kind != JVal::STR)
                fatal(AMGX_RC_BAD_CONFIGURATION, "JSON object \"" + name + "\" has no \"solver\" string member");
            JVal tmp; tmp.kind = JVal::STR; tmp.s = sv->s;
            imp(cfg, name, tmp, cur, sub);
            // children use `sub` as their current scope: emulate by injecting scope
            JVal child = v;
            bool has = false;
            for (auto &cm : child.members) if (cm.first == "scope") { has = true; }
            if (!has) { JVal sc; sc.kind = JVal::STR; sc.s = sub; child.members.emplace_back("scope", sc); }
            import_json(cfg, child, false, imp);
        } else if (v.kind == JVal::INT || v.kind == JVal::DBL || v.kind == JVal::STR) {
            imp(cfg, name, v, cur, "default");
        } else if (v.kind == JVal::BOOL) {
            JVal t; t.kind = JVal::INT; t.i = v.b ? 1 : 0; t.raw = v.b ? "1" : "0";
            imp(cfg, name, t, cur, "default");
        }
        // arrays / null: ignored like the reference (it builds an error string and drops it)
    }
}

bool Config::parse_json(const char *str)
{
    JParser jp{str, str + strlen(str)};
    JVal root;
    if (!jp.value(root) || root.kind != JVal::OBJ) return false;
    jp.ws();
    if (jp.p != jp.end) return false;
    auto imp = [](Config &c, const std::string &name, const JVal &v, const std::string &cur, const std::string &ns) {
        const ParamDesc *d = Config::find_desc(name);
        if (!d) fatal(AMGX_RC_BAD_CONFIGURATION, "Variable '" + name + "' not registered");
        if (v.kind == JVal::STR) {
            c.import_named(name, v.s, true, false, cur, ns);
        } else if (v.kind == JVal::INT) {
            if (d->type == PType::STRING)
                fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry. Type of the parameter \"" + name + "\" in the config is int, but string is expected");
            c.import_named(name, v.raw, false, false, cur, ns);
        } else {   // double; ints registered -> truncation like (int)(c_value)
            if (d->type == PType::STRING)
                fatal(AMGX_RC_BAD_CONFIGURATION, "Incorrect config entry. Type of the parameter \"" + name + "\" in the config is double, but string is expected");
            if (d->type == PType::DOUBLE) c.import_named(name, v.raw, false, true, cur, ns);
            else c.import_named(name, std::to_string((long long)v.d), false, true, cur, ns);
        }
    };
    import_json(*this, root, true, imp);
    return true;
}

void Config::parse_string(const char *str)
{
    if (!str) fatal(AMGX_RC_BAD_CONFIGURATION, "NULL configuration string");
    // JSON first
    const char *q = str;
    while (*q && isspace((unsigned char)*q)) q++;
    if (*q == '{') {
        if (parse_json(str)) return;
        fatal(AMGX_RC_BAD_CONFIGURATION, "Cannot parse configuration as JSON");
    }
    parse_legacy(str);
}

void Config::parse_file(const char *filename)
{
    std::ifstream fin(filename);
    if (!fin) fatal(AMGX_RC_IO_ERROR, std::string("Error: Cannot read config file: ") + (filename ? filename : "(null)"));
    std::stringstream ss;
    ss << fin.rdbuf();
    std::string content = ss.str();
    size_t k = 0;
    while (k < content.size() && isspace((unsigned char)content[k])) k++;
    if (k < content.size() && content[k] == '{') {
        if (!parse_json(content.c_str()))
            fatal(AMGX_RC_BAD_CONFIGURATION, std::string("Error: Cannot import config from JSON file: ") + filename);
        return;
    }
    // legacy file: one entry per line, '#' comments (src/amg_config.cu:327-371)
    std::string params, line;
    std::istringstream is(content);
    while (std::getline(is, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        params += line + ", ";
    }
    parse_legacy(params);
}

const ParamValue *Config::lookup(const std::string &name, const std::string &scope, const ParamDesc **d) const
{
    *d = find_desc(name);
    if (!*d) fatal(AMGX_RC_BAD_CONFIGURATION, "getParameter error: '" + name + "' not found");
    auto it = params_.find({scope, name});
    return it == params_.end() ? nullptr : &it->second;
}

int Config::get_int(const std::string &name, const std::string &scope) const
{
    const ParamDesc *d;
    const ParamValue *v = lookup(name, scope, &d);
    if (d->type != PType::INT && d->type != PType::SIZE) fatal(AMGX_RC_BAD_CONFIGURATION, "getParameter error: '" + name + "' type miss match");
    return v ? (int)v->i : (int)atoll(d->def);
}
double Config::get_double(const std::string &name, const std::string &scope) const
{
    const ParamDesc *d;
    const ParamValue *v = lookup(name, scope, &d);
    if (d->type != PType::DOUBLE) fatal(AMGX_RC_BAD_CONFIGURATION, "getParameter error: '" + name + "' type miss match");
    return v ? v->d : strtod(d->def, nullptr);
}
std::string Config::get_string(const std::string &name, const std::string &scope) const
{
    const ParamDesc *d;
    const ParamValue *v = lookup(name, scope, &d);
    if (d->type != PType::STRING) fatal(AMGX_RC_BAD_CONFIGURATION, "getParameter error: '" + name + "' type miss match");
    return v ? v->s : std::string(d->def);
}
void Config::get_scoped(const std::string &name, const std::string &scope, std::string &value, std::string &new_scope) const
{
    const ParamDesc *d;
    const ParamValue *v = lookup(name, scope, &d);
    if (d->type != PType::STRING) fatal(AMGX_RC_BAD_CONFIGURATION, "getParameter error: '" + name + "' type miss match");
    if (v) { value = v->s; new_scope = v->new_scope; }
    else { value = d->def; new_scope = "default"; }
}
bool Config::is_set(const std::string &name, const std::string &scope) const
{
    return params_.find({scope, name}) != params_.end();
}
void Config::set_int(const std::string &name, long long val, const std::string &scope)
{
    const ParamDesc *d = find_desc(name);
    if (!d) fatal(AMGX_RC_BAD_CONFIGURATION, "setParameter error: '" + name + "' not found");
    ParamValue v; v.type = d->type; v.i = val;
    params_[{scope, name}] = v;
}

}  // namespace amgxb
