// raptor_b200/csrc/json_io.cu -- the reference's parameter / state JSON wire format on the flat rows of include/b200_l2f.h (host code only).
//
// Export follows rl_tools::json (rl/environments/l2f/operations_cpu.h:139-411 parameters, :412-560 state) character for character: the same
// key order and separators and std::to_string formatting ("%f" of the value promoted to double).  Import follows rl_tools::from_json
// (:565-824): every key the reference reads is required, numbers are parsed as double and narrowed to float (what nlohmann::json does for
// `float x = j[...]`), unknown keys are ignored, booleans must be JSON booleans.  Two behaviours of the reference are kept on purpose:
//   * the state writer / reader never reach the random-force layer (its overloads are declared after the rotor layers that would call them,
//     and are not found by argument-dependent lookup), so "force" / "torque" are neither written nor read;
//   * "current_step" of the action-history ring is not part of the wire format.
// The parser is a small recursive-descent JSON reader (objects, arrays, numbers, strings, true / false / null); the reference uses nlohmann::json,
// an un-vendored submodule (rl-tools/external/json).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "handle.h"

using namespace b200l2f;

namespace {

struct JValue {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    bool b = false; double num = 0.0; std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;
    const JValue* get(const char* key) const {
        if(kind != OBJ) return nullptr;
        for(auto& kv : obj) if(kv.first == key) return &kv.second;
        return nullptr;
    }
};

struct Parser {
    const char* p; const char* end; std::string err;
    void ws(){ while(p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool fail(const std::string& m){ if(err.empty()) err = m; return false; }
    bool parse_string(std::string& out){
        if(p >= end || *p != '"') return fail("expected string");
        p++;
        while(p < end && *p != '"'){
            if(*p == '\\'){
                if(++p >= end) return fail("bad escape");
                switch(*p){
                    case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; case 'b': out += '\b'; break; case 'f': out += '\f'; break;
                    case 'u': { if(end - p < 5) return fail("bad \\u escape"); unsigned v = (unsigned)std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16); out += (char)(v < 128 ? v : '?'); p += 4; break; }
                    default: out += *p;
                }
                p++;
            }
            else out += *p++;
        }
        if(p >= end) return fail("unterminated string");
        p++;
        return true;
    }
    bool parse_value(JValue& v, int depth){
        if(depth > 32) return fail("nesting too deep");
        ws();
        if(p >= end) return fail("unexpected end of input");
        if(*p == '{'){
            v.kind = JValue::OBJ; p++; ws();
            if(p < end && *p == '}'){ p++; return true; }
            for(;;){
                ws();
                std::string key;
                if(!parse_string(key)) return false;
                ws();
                if(p >= end || *p != ':') return fail("expected ':'");
                p++;
                JValue child;
                if(!parse_value(child, depth + 1)) return false;
                v.obj.emplace_back(std::move(key), std::move(child));
                ws();
                if(p < end && *p == ','){ p++; continue; }
                if(p < end && *p == '}'){ p++; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if(*p == '['){
            v.kind = JValue::ARR; p++; ws();
            if(p < end && *p == ']'){ p++; return true; }
            for(;;){
                JValue child;
                if(!parse_value(child, depth + 1)) return false;
                v.arr.push_back(std::move(child));
                ws();
                if(p < end && *p == ','){ p++; continue; }
                if(p < end && *p == ']'){ p++; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if(*p == '"'){ v.kind = JValue::STR; return parse_string(v.str); }
        if(end - p >= 4 && !std::strncmp(p, "true", 4)){ v.kind = JValue::BOOL; v.b = true; p += 4; return true; }
        if(end - p >= 5 && !std::strncmp(p, "false", 5)){ v.kind = JValue::BOOL; v.b = false; p += 5; return true; }
        if(end - p >= 4 && !std::strncmp(p, "null", 4)){ v.kind = JValue::NUL; p += 4; return true; }
        char* stop = nullptr;
        const double d = std::strtod(p, &stop);
        if(stop == p) return fail("unexpected character");
        v.kind = JValue::NUM; v.num = d; p = stop;
        return true;
    }
};

// path-tracking reader over the parsed tree: the first missing / mistyped key is reported with its path
struct Reader {
    std::string err;
    const JValue* child(const JValue* o, const char* key, const std::string& path){
        if(!o) return nullptr;
        const JValue* c = o->get(key);
        if(!c && err.empty()) err = "missing key \"" + path + key + "\"";
        return c;
    }
    const JValue* at(const JValue* a, size_t i, const std::string& path){
        if(!a) return nullptr;
        if(a->kind != JValue::ARR || i >= a->arr.size()){ if(err.empty()) err = "\"" + path + "\": expected an array with more than " + std::to_string(i) + " entries"; return nullptr; }
        return &a->arr[i];
    }
    void num(const JValue* v, float& out, const std::string& path){
        if(!v) return;
        if(v->kind == JValue::NUM) out = (float)v->num;
        else if(v->kind == JValue::BOOL) out = v->b ? 1.0f : 0.0f;      // nlohmann converts booleans to arithmetic types as well
        else if(err.empty()) err = "\"" + path + "\": expected a number";
    }
    void boolean(const JValue* v, float& out, const std::string& path){
        if(!v) return;
        if(v->kind == JValue::BOOL) out = v->b ? 1.0f : 0.0f;
        else if(err.empty()) err = "\"" + path + "\": expected true / false";
    }
    void scalar(const JValue* o, const char* key, float& out, const std::string& path){ num(child(o, key, path), out, path + key); }
    void flag(const JValue* o, const char* key, float& out, const std::string& path){ boolean(child(o, key, path), out, path + key); }
    void vec(const JValue* o, const char* key, float* out, int n, const std::string& path){
        const JValue* a = child(o, key, path);
        for(int i = 0; i < n; i++) num(at(a, i, path + key), out[i], path + key);
    }
    void mat(const JValue* o, const char* key, float* out, int rows, int cols, const std::string& path){
        const JValue* a = child(o, key, path);
        for(int r = 0; r < rows; r++){ const JValue* row = at(a, r, path + key); for(int c = 0; c < cols; c++) num(at(row, c, path + key), out[r * cols + c], path + key); }
    }
};

std::string f2s(float v){ char b[64]; std::snprintf(b, sizeof(b), "%f", (double)v); return b; }   // std::to_string(float)
std::string vec2s(const float* v, int n){ std::string s = "["; for(int i = 0; i < n; i++){ s += f2s(v[i]); if(i < n - 1) s += ", "; } return s + "]"; }
std::string mat2s(const float* v, int rows, int cols){ std::string s = "["; for(int r = 0; r < rows; r++){ s += vec2s(v + r * cols, cols); if(r < rows - 1) s += ", "; } return s + "]"; }
const char* b2s(float v){ return v != 0.0f ? "true" : "false"; }

std::string parameters_json(const float* p){
    std::string s = "{\"dynamics\": {";
    s += "\"rotor_positions\": " + mat2s(p + P_ROTOR_POS, 4, 3) + ", ";
    s += "\"rotor_thrust_directions\": " + mat2s(p + P_THRUST_DIR, 4, 3) + ", ";
    s += "\"rotor_torque_directions\": " + mat2s(p + P_TORQUE_DIR, 4, 3) + ", ";
    s += "\"rotor_thrust_coefficients\": " + mat2s(p + P_THRUST_COEF, 4, 3) + ", ";
    s += "\"rotor_torque_constants\": " + vec2s(p + P_TORQUE_CONST, 4) + ", ";
    s += "\"rotor_time_constants_rising\": " + vec2s(p + P_TAU_RISE, 4) + ", ";
    s += "\"rotor_time_constants_falling\": " + vec2s(p + P_TAU_FALL, 4) + ", ";
    s += "\"mass\": " + f2s(p[P_MASS]) + ", ";
    s += "\"gravity\": " + vec2s(p + P_GRAVITY, 3) + ", ";
    s += "\"J\": " + mat2s(p + P_J, 3, 3) + ", ";
    s += "\"J_inv\": " + mat2s(p + P_JINV, 3, 3) + ", ";
    s += "\"hovering_throttle_relative\": " + f2s(p[P_HOVER]) + ", ";
    s += "\"action_limit\": {\"min\": " + f2s(p[P_ACT_MIN]) + ", \"max\": " + f2s(p[P_ACT_MAX]) + "}}, ";
    s += "\"integration\": {\"dt\": " + f2s(p[P_DT]) + "}, ";
    s += "\"mdp\": {\"init\": {";
    s += "\"guidance\": " + f2s(p[P_INIT_GUIDANCE]) + ", \"max_position\": " + f2s(p[P_INIT_MAX_POS]) + ", \"max_angle\": " + f2s(p[P_INIT_MAX_ANGLE]) + ", ";
    s += "\"max_linear_velocity\": " + f2s(p[P_INIT_MAX_LINVEL]) + ", \"max_angular_velocity\": " + f2s(p[P_INIT_MAX_ANGVEL]) + ", ";
    s += std::string("\"relative_rpm\": ") + b2s(p[P_INIT_REL_RPM]) + ", \"min_rpm\": " + f2s(p[P_INIT_MIN_RPM]) + ", \"max_rpm\": " + f2s(p[P_INIT_MAX_RPM]) + "}, ";
    s += std::string("\"reward\": {\"non_negative\": ") + b2s(p[P_RW_NONNEG]) + ", \"scale\": " + f2s(p[P_RW_SCALE]) + ", \"constant\": " + f2s(p[P_RW_CONSTANT]) + ", ";
    s += "\"termination_penalty\": " + f2s(p[P_RW_TERM_PENALTY]) + ", \"position\": " + f2s(p[P_RW_POSITION]) + ", \"position_clip\": " + f2s(p[P_RW_POSITION_CLIP]) + ", ";
    s += "\"orientation\": " + f2s(p[P_RW_ORIENTATION]) + ", \"linear_velocity\": " + f2s(p[P_RW_LINVEL]) + ", \"angular_velocity\": " + f2s(p[P_RW_ANGVEL]) + ", ";
    s += "\"linear_acceleration\": " + f2s(p[P_RW_LINACC]) + ", \"angular_acceleration\": " + f2s(p[P_RW_ANGACC]) + ", \"action\": " + f2s(p[P_RW_ACTION]) + ", ";
    s += "\"d_action\": " + f2s(p[P_RW_DACTION]) + ", \"position_error_integral\": " + f2s(p[P_RW_POS_INTEGRAL]) + "}, ";
    s += "\"observation_noise\": {\"position\": " + f2s(p[P_NOISE_POS]) + ", \"orientation\": " + f2s(p[P_NOISE_ORI]) + ", \"linear_velocity\": " + f2s(p[P_NOISE_LINVEL]) + ", ";
    s += "\"angular_velocity\": " + f2s(p[P_NOISE_ANGVEL]) + ", \"imu_acceleration\": " + f2s(p[P_NOISE_IMU]) + "}, ";
    s += "\"action_noise\": {\"normalized_rpm\": " + f2s(p[P_ACTION_NOISE]) + "}, ";
    s += std::string("\"termination\": {\"enabled\": ") + b2s(p[P_TERM_ENABLED]) + ", \"position_threshold\": " + f2s(p[P_TERM_POS]) + ", \"linear_velocity_threshold\": " + f2s(p[P_TERM_LINVEL]) + ", ";
    s += "\"angular_velocity_threshold\": " + f2s(p[P_TERM_ANGVEL]) + ", \"position_integral_threshold\": " + f2s(p[P_TERM_POS_INT]) + ", \"orientation_integral_threshold\": " + f2s(p[P_TERM_ORI_INT]) + "}}, ";
    s += "\"disturbances\": {\"random_force\": {\"mean\": " + f2s(p[P_DIST_FORCE_MEAN]) + ", \"std\": " + f2s(p[P_DIST_FORCE_STD]) + "}, ";
    s += "\"random_torque\": {\"mean\": " + f2s(p[P_DIST_TORQUE_MEAN]) + ", \"std\": " + f2s(p[P_DIST_TORQUE_STD]) + "}}, ";
    static const char* dr_keys[15] = {"thrust_to_weight_min", "thrust_to_weight_max", "torque_to_inertia_min", "torque_to_inertia_max", "mass_min", "mass_max", "mass_size_deviation",
        "rotor_time_constant_rising_min", "rotor_time_constant_rising_max", "rotor_time_constant_falling_min", "rotor_time_constant_falling_max", "rotor_torque_constant_min",
        "rotor_torque_constant_max", "orientation_offset_angle_max", "disturbance_force_max"};
    s += "\"domain_randomization\": {";
    for(int i = 0; i < 15; i++){ s += std::string("\"") + dr_keys[i] + "\": " + f2s(p[P_DR_T2W_MIN + i]); if(i < 14) s += ", "; }
    s += "}, ";
    s += "\"trajectory\": {\"MIXTURE_N\": 2, \"mixture\": " + vec2s(p + P_TRAJ_MIX0, 2) + ", ";
    s += "\"langevin\": {\"gamma\": " + f2s(p[P_LANGEVIN_GAMMA]) + ", \"omega\": " + f2s(p[P_LANGEVIN_OMEGA]) + ", \"sigma\": " + f2s(p[P_LANGEVIN_SIGMA]) + ", \"alpha\": " + f2s(p[P_LANGEVIN_ALPHA]) + "}}}";
    return s;
}

std::string parameters_from(const JValue& root, float* p){
    Reader r;
    const JValue* dyn = r.child(&root, "dynamics", "");
    r.mat(dyn, "rotor_positions", p + P_ROTOR_POS, 4, 3, "dynamics."); r.mat(dyn, "rotor_thrust_directions", p + P_THRUST_DIR, 4, 3, "dynamics.");
    r.mat(dyn, "rotor_torque_directions", p + P_TORQUE_DIR, 4, 3, "dynamics."); r.mat(dyn, "rotor_thrust_coefficients", p + P_THRUST_COEF, 4, 3, "dynamics.");
    r.vec(dyn, "rotor_torque_constants", p + P_TORQUE_CONST, 4, "dynamics."); r.vec(dyn, "rotor_time_constants_rising", p + P_TAU_RISE, 4, "dynamics.");
    r.vec(dyn, "rotor_time_constants_falling", p + P_TAU_FALL, 4, "dynamics.");
    r.scalar(dyn, "mass", p[P_MASS], "dynamics."); r.vec(dyn, "gravity", p + P_GRAVITY, 3, "dynamics.");
    r.mat(dyn, "J", p + P_J, 3, 3, "dynamics."); r.mat(dyn, "J_inv", p + P_JINV, 3, 3, "dynamics.");
    r.scalar(dyn, "hovering_throttle_relative", p[P_HOVER], "dynamics.");
    const JValue* lim = r.child(dyn, "action_limit", "dynamics.");
    r.scalar(lim, "min", p[P_ACT_MIN], "dynamics.action_limit."); r.scalar(lim, "max", p[P_ACT_MAX], "dynamics.action_limit.");
    r.scalar(r.child(&root, "integration", ""), "dt", p[P_DT], "integration.");
    const JValue* mdp = r.child(&root, "mdp", "");
    const JValue* init = r.child(mdp, "init", "mdp.");
    r.scalar(init, "guidance", p[P_INIT_GUIDANCE], "mdp.init."); r.scalar(init, "max_position", p[P_INIT_MAX_POS], "mdp.init."); r.scalar(init, "max_angle", p[P_INIT_MAX_ANGLE], "mdp.init.");
    r.scalar(init, "max_linear_velocity", p[P_INIT_MAX_LINVEL], "mdp.init."); r.scalar(init, "max_angular_velocity", p[P_INIT_MAX_ANGVEL], "mdp.init.");
    r.flag(init, "relative_rpm", p[P_INIT_REL_RPM], "mdp.init."); r.scalar(init, "min_rpm", p[P_INIT_MIN_RPM], "mdp.init."); r.scalar(init, "max_rpm", p[P_INIT_MAX_RPM], "mdp.init.");
    const JValue* rw = r.child(mdp, "reward", "mdp.");
    r.flag(rw, "non_negative", p[P_RW_NONNEG], "mdp.reward.");
    static const char* rw_keys[13] = {"scale", "constant", "termination_penalty", "position", "position_clip", "orientation", "linear_velocity", "angular_velocity", "linear_acceleration",
        "angular_acceleration", "action", "d_action", "position_error_integral"};
    for(int i = 0; i < 13; i++) r.scalar(rw, rw_keys[i], p[P_RW_SCALE + i], "mdp.reward.");
    const JValue* on = r.child(mdp, "observation_noise", "mdp.");
    static const char* on_keys[5] = {"position", "orientation", "linear_velocity", "angular_velocity", "imu_acceleration"};
    for(int i = 0; i < 5; i++) r.scalar(on, on_keys[i], p[P_NOISE_POS + i], "mdp.observation_noise.");
    r.scalar(r.child(mdp, "action_noise", "mdp."), "normalized_rpm", p[P_ACTION_NOISE], "mdp.action_noise.");
    const JValue* term = r.child(mdp, "termination", "mdp.");
    r.flag(term, "enabled", p[P_TERM_ENABLED], "mdp.termination.");
    static const char* term_keys[5] = {"position_threshold", "linear_velocity_threshold", "angular_velocity_threshold", "position_integral_threshold", "orientation_integral_threshold"};
    for(int i = 0; i < 5; i++) r.scalar(term, term_keys[i], p[P_TERM_POS + i], "mdp.termination.");
    const JValue* dist = r.child(&root, "disturbances", "");
    const JValue* rf = r.child(dist, "random_force", "disturbances."); const JValue* rt = r.child(dist, "random_torque", "disturbances.");
    r.scalar(rf, "mean", p[P_DIST_FORCE_MEAN], "disturbances.random_force."); r.scalar(rf, "std", p[P_DIST_FORCE_STD], "disturbances.random_force.");
    r.scalar(rt, "mean", p[P_DIST_TORQUE_MEAN], "disturbances.random_torque."); r.scalar(rt, "std", p[P_DIST_TORQUE_STD], "disturbances.random_torque.");
    const JValue* dr = r.child(&root, "domain_randomization", "");
    static const char* dr_keys[15] = {"thrust_to_weight_min", "thrust_to_weight_max", "torque_to_inertia_min", "torque_to_inertia_max", "mass_min", "mass_max", "mass_size_deviation",
        "rotor_time_constant_rising_min", "rotor_time_constant_rising_max", "rotor_time_constant_falling_min", "rotor_time_constant_falling_max", "rotor_torque_constant_min",
        "rotor_torque_constant_max", "orientation_offset_angle_max", "disturbance_force_max"};
    for(int i = 0; i < 15; i++) r.scalar(dr, dr_keys[i], p[P_DR_T2W_MIN + i], "domain_randomization.");
    const JValue* traj = r.child(&root, "trajectory", "");
    float mixture_n = 0.0f;
    r.scalar(traj, "MIXTURE_N", mixture_n, "trajectory.");
    if(r.err.empty() && mixture_n != 2.0f) r.err = "Mismatch in MIXTURE_N";                       // operations_cpu.h:708
    r.vec(traj, "mixture", p + P_TRAJ_MIX0, 2, "trajectory.");
    const JValue* lg = r.child(traj, "langevin", "trajectory.");
    r.scalar(lg, "gamma", p[P_LANGEVIN_GAMMA], "trajectory.langevin."); r.scalar(lg, "omega", p[P_LANGEVIN_OMEGA], "trajectory.langevin.");
    r.scalar(lg, "sigma", p[P_LANGEVIN_SIGMA], "trajectory.langevin."); r.scalar(lg, "alpha", p[P_LANGEVIN_ALPHA], "trajectory.langevin.");
    return r.err;
}

std::string state_json(const float* s, int H, bool langevin_spec){
    std::string j = "{";
    j += "\"position\": " + vec2s(s + S_POS, 3) + ", \"orientation\": " + vec2s(s + S_ORI, 4) + ", \"linear_velocity\": " + vec2s(s + S_LINVEL, 3) + ", \"angular_velocity\": " + vec2s(s + S_ANGVEL, 3) + ", ";
    j += "\"last_action\": " + vec2s(s + S_LAST_ACTION, 4) + ", ";
    j += "\"angular_velocity_history\": " + mat2s(s + S_ANGVEL_HIST, 1, 3) + ", ";
    j += "\"rpm\": " + vec2s(s + S_RPM, 4) + ", ";
    j += "\"action_history\": " + mat2s(s + S_HIST, H, 4) + ", ";
    j += "\"trajectory\": {\"type\": ";
    const int type = (int)s[s_traj_type(H)];
    if(type == 0) j += "\"POSITION\"";
    else if(type == 1){
        const float* l = s + s_langevin(H);
        j += "\"LANGEVIN\", \"langevin\": {\"position\": " + vec2s(l, 3) + ", \"velocity\": " + vec2s(l + 3, 3) + ", \"position_raw\": " + vec2s(l + 6, 3) + ", \"velocity_raw\": " + vec2s(l + 9, 3) + "}";
    }
    else j += "\"NONE\"";
    (void)langevin_spec;
    return j + "}}";
}
std::string state_from(const JValue& root, float* s, int H){
    Reader r;
    r.vec(&root, "position", s + S_POS, 3, ""); r.vec(&root, "orientation", s + S_ORI, 4, ""); r.vec(&root, "linear_velocity", s + S_LINVEL, 3, ""); r.vec(&root, "angular_velocity", s + S_ANGVEL, 3, "");
    r.vec(&root, "last_action", s + S_LAST_ACTION, 4, "");
    // StateAngularVelocityDelay<0>: the reader loops over HISTORY_LENGTH = 0 entries (operations_cpu.h:760-764), the single memory slot stays as it is
    r.vec(&root, "rpm", s + S_RPM, 4, "");
    r.mat(&root, "action_history", s + S_HIST, H, 4, "");
    const JValue* traj = r.child(&root, "trajectory", "");
    const JValue* type = r.child(traj, "type", "trajectory.");
    if(type){
        if(type->kind != JValue::STR){ if(r.err.empty()) r.err = "\"trajectory.type\": expected a string"; }
        else if(type->str == "POSITION") s[s_traj_type(H)] = 0.0f;
        else if(type->str == "LANGEVIN"){
            s[s_traj_type(H)] = 1.0f;
            const JValue* l = r.child(traj, "langevin", "trajectory.");
            float* d = s + s_langevin(H);
            r.vec(l, "position", d, 3, "trajectory.langevin."); r.vec(l, "velocity", d + 3, 3, "trajectory.langevin.");
            r.vec(l, "position_raw", d + 6, 3, "trajectory.langevin."); r.vec(l, "velocity_raw", d + 9, 3, "trajectory.langevin.");
        }
    }
    return r.err;
}

int emit(b200l2f_handle* h, const std::string& s, char* buf, size_t capacity, size_t* length){
    if(length) *length = s.size();
    if(!buf || capacity < s.size() + 1) return fail(h, B200L2F_ERR_ARGUMENT, "json: buffer too small (the required size without the terminator is returned in *length)");
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return B200L2F_OK;
}
int parse(b200l2f_handle* h, const char* json, JValue& root){
    if(!json) return fail(h, B200L2F_ERR_ARGUMENT, "json: null string");
    Parser ps{json, json + std::strlen(json), {}};
    if(!ps.parse_value(root, 0)) return fail(h, B200L2F_ERR_ARGUMENT, "json: parse error at offset " + std::to_string(ps.p - json) + ": " + ps.err);
    ps.ws();
    if(ps.p != ps.end) return fail(h, B200L2F_ERR_ARGUMENT, "json: trailing characters at offset " + std::to_string(ps.p - json));
    return B200L2F_OK;
}

}  // namespace

extern "C" {

int b200l2f_parameters_to_json(b200l2f_handle* h, const float* row145, char* buf, size_t capacity, size_t* length){
    if(!row145) return fail(h, B200L2F_ERR_ARGUMENT, "parameters_to_json: null row");
    return emit(h, parameters_json(row145), buf, capacity, length);
}
int b200l2f_parameters_from_json(b200l2f_handle* h, const char* json, float* row145_io){
    if(!row145_io) return fail(h, B200L2F_ERR_ARGUMENT, "parameters_from_json: null row");
    JValue root; int rc;
    if((rc = parse(h, json, root))) return rc;
    float tmp[B200L2F_PARAMS_DIM];
    std::memcpy(tmp, row145_io, sizeof(tmp));
    const std::string err = parameters_from(root, tmp);
    if(!err.empty()) return fail(h, B200L2F_ERR_ARGUMENT, "parameters_from_json: " + err);
    std::memcpy(row145_io, tmp, sizeof(tmp));
    return B200L2F_OK;
}
int b200l2f_state_to_json(b200l2f_handle* h, const float* state_row, char* buf, size_t capacity, size_t* length){
    if(!h || !state_row) return fail(h, B200L2F_ERR_ARGUMENT, "state_to_json: null argument");
    return emit(h, state_json(state_row, h->H, h->kind != KIND_DEFAULT), buf, capacity, length);
}
int b200l2f_state_from_json(b200l2f_handle* h, const char* json, float* state_row_io){
    if(!h || !state_row_io) return fail(h, B200L2F_ERR_ARGUMENT, "state_from_json: null argument");
    JValue root; int rc;
    if((rc = parse(h, json, root))) return rc;
    std::vector<float> tmp(state_row_io, state_row_io + h->sdim);
    const std::string err = state_from(root, tmp.data(), h->H);
    if(!err.empty()) return fail(h, B200L2F_ERR_ARGUMENT, "state_from_json: " + err);
    std::memcpy(state_row_io, tmp.data(), sizeof(float) * h->sdim);
    return B200L2F_OK;
}

}  // extern "C"
