// config.h -- configuration surface of the AMGX C API: JSON files / JSON strings / legacy
// "scope:name(new_scope)=value" strings, stored per (scope, name) with registry defaults.
// Behaviour follows the reference's AMG_Config (src/amg_config.cu:60-330 legacy parser,
// :545-700 JSON import, :980-1010 lookup with default fallback; registry src/core.cu:307-543).
#pragma once
#include "base.h"
#include <map>

namespace amgxb {

enum class PType { INT, SIZE, DOUBLE, STRING };

struct ParamDesc {
    const char *name;
    PType type;
    const char *def;      // default value, textual
};

struct ParamValue {
    PType type = PType::INT;
    long long i = 0;
    double d = 0.0;
    std::string s;
    std::string new_scope = "default";
};

class Config {
public:
    Config() = default;
    // Throws Error(AMGX_RC_BAD_CONFIGURATION / AMGX_RC_IO_ERROR)
    void parse_string(const char *str);      // JSON first, then legacy format
    void parse_file(const char *filename);

    int         get_int(const std::string &name, const std::string &scope) const;
    double      get_double(const std::string &name, const std::string &scope) const;
    std::string get_string(const std::string &name, const std::string &scope) const;
    // value + the scope the named sub-solver lives in (for solver/preconditioner/smoother/coarse_solver)
    void get_scoped(const std::string &name, const std::string &scope, std::string &value,
                    std::string &new_scope) const;
    bool is_set(const std::string &name, const std::string &scope) const;
    void set_int(const std::string &name, long long v, const std::string &scope);

    static const ParamDesc *find_desc(const std::string &name);
    static const ParamDesc *registry(size_t *count);

    bool allow_mod = false;   // AMGX_config_add_parameters sets this while adding

private:
    std::map<std::pair<std::string, std::string>, ParamValue> params_;
    std::vector<std::string> scopes_{"default"};

    const ParamValue *lookup(const std::string &name, const std::string &scope, const ParamDesc **d) const;
    void import_named(const std::string &name, const std::string &textual, bool is_string_token,
                      bool is_double_token, const std::string &cur_scope, const std::string &new_scope);
    void parse_legacy(std::string params);
    bool parse_json(const char *str);   // returns false when `str` is not JSON at all
    void set_one_legacy(const std::string &entry);
};

}  // namespace amgxb
