// raptor_b200/csrc/checkpoint_io.cu -- reader for rl-tools' checkpoint CODE EXPORT (`checkpoint.h`), host code only.
//
// The reference stores trained actors as a generated C++ header: rl::loop::steps::checkpoint::save_code
// (rl/loop/steps/checkpoint/operations_cpu.h:56-118) writes
//     rl_tools::save_code(device, actor, "rl_tools::checkpoint::actor", true)   -- nested namespaces, one per layer / parameter, each parameter as
//                                                                                  `alignas(T) const unsigned char memory[] = {b0, b1, ...};`
//                                                                                  (containers/{matrix,tensor}/persist_code.h) + shape aliases
//     ... "rl_tools::checkpoint::example::input" / "::output"                   -- a known-answer pair (randn input, Evaluation-mode output)
//     namespace rl_tools::checkpoint::meta{ char name[] = "..."; char commit_hash[] = "..."; }
// and the reference consumes it by COMPILING it in (post_training/load_actor.cpp, inference/applications/l2f/c_backend.h).  An engine behind
// a C ABI cannot compile headers at run time, so this unit reads the same text: a scanner that follows the namespace nesting, decodes every
// `memory[]` byte list into floats with the shape declared next to it, and keeps the CONFIG / INPUT_SHAPE aliases of every namespace.  On top of
// the tensor list, b200l2f_checkpoint_policy recognises the actor architectures the engine runs and assembles their weight blob in the order
// include/b200_l2f.h documents:
//     Sequential<Dense(ReLU), GRU, Dense>                         (the Raptor checkpoint, checkpoint.h:40-185)       -> B200L2F_POLICY_RAPTOR_GRU
//     [Standardize ->] MLP(3 layers, ReLU) [-> SampleAndSquash]   (SAC teachers, rl/algorithms/sac/loop/core/approximators_mlp.h:14-37;
//                                                                   PPO actors with mlp_unconditional_stddev: log_std)  -> B200L2F_POLICY_MLP
// The HDF5 twin (checkpoint.h5, operations_cpu.h:119-160; the reference reads it through HighFive / libhdf5, which this image does not have) is
// read by the engine's own minimal HDF5 reader (h5_io.cu) and mapped onto the same tensor paths (from_h5 below), so both files of a checkpoint
// folder give the same tensors and the same policy blob.
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "handle.h"
#include "h5_io.h"

using namespace b200l2f;

struct b200l2f_checkpoint {
    struct TensorEntry { std::string path; std::vector<int64_t> dims; std::vector<float> data; int64_t bytes = 0; int elem_size = 4; int64_t row_alignment = 1; bool shaped = false; };
    std::vector<TensorEntry> tensors;
    std::map<std::string, std::string> config;                 // namespace path -> text of `using CONFIG = ...`
    std::map<std::string, std::vector<int64_t>> input_shape;   // namespace path -> dims of `using INPUT_SHAPE = ...Shape<TI, ...>`
    std::map<std::string, std::string> strings;                // namespace path + "::" + name -> value of `char name[] = "..."`
    std::string err;
    const TensorEntry* find(const std::string& path) const { for(auto& t : tensors) if(t.path == path) return &t; return nullptr; }
};

namespace {

using Ckpt = b200l2f_checkpoint;

struct Scanner {
    const char* p; const char* end; Ckpt& out;
    struct Frame { int names; };            // how many path components this brace pushed (0 = anonymous block)
    std::vector<Frame> frames;
    std::vector<std::string> path;
    int pending_elem_size = 4;              // from the `alignas(T)` that precedes a memory[] declaration

    std::string joined() const { std::string s; for(size_t i = 0; i < path.size(); i++){ if(i) s += "::"; s += path[i]; } return s; }
    // path without the storage namespace the reference wraps parameter memory in (nn/parameters/persist_code.h)
    std::string tensor_path() const {
        std::string s;
        for(auto& c : path){ if(c == "parameters_memory") continue; if(!s.empty()) s += "::"; s += c; }
        return s;
    }
    bool fail(const std::string& m){ if(out.err.empty()) out.err = m + " (offset " + std::to_string((long long)(p - begin)) + ")"; return false; }
    const char* begin;

    void skip_ws(){
        for(;;){
            while(p < end && std::isspace((unsigned char)*p)) p++;
            if(p + 1 < end && p[0] == '/' && p[1] == '/'){ while(p < end && *p != '\n') p++; continue; }
            if(p + 1 < end && p[0] == '/' && p[1] == '*'){ p += 2; while(p + 1 < end && !(p[0] == '*' && p[1] == '/')) p++; p = p + 2 <= end ? p + 2 : end; continue; }
            if(p < end && *p == '#'){ while(p < end && *p != '\n') p++; continue; }   // #include lines of the export
            break;
        }
    }
    static bool ident_char(char c){ return std::isalnum((unsigned char)c) || c == '_'; }
    std::string ident(){ const char* s = p; while(p < end && ident_char(*p)) p++; return std::string(s, p); }
    bool skip_string(std::string* value){
        const char q = *p++;
        std::string v;
        while(p < end && *p != q){ if(*p == '\\' && p + 1 < end){ p++; } v += *p++; }
        if(p >= end) return fail("unterminated literal");
        p++;
        if(value) *value = v;
        return true;
    }
    // the integers of `Name<a, b, c<d, e>, f>` starting at the '<' p points to; nested template arguments are skipped, non-numeric ones ignored
    static std::vector<int64_t> template_integers(const std::string& text, size_t lt){
        std::vector<int64_t> v;
        int depth = 0; size_t i = lt;
        std::string tok;
        auto flush = [&](){
            size_t a = 0; while(a < tok.size() && std::isspace((unsigned char)tok[a])) a++;
            size_t b = tok.size(); while(b > a && std::isspace((unsigned char)tok[b - 1])) b--;
            if(b > a){ bool num = true; for(size_t k = a; k < b; k++) num = num && std::isdigit((unsigned char)tok[k]); if(num) v.push_back(std::strtoll(tok.substr(a, b - a).c_str(), nullptr, 10)); }
            tok.clear();
        };
        for(; i < text.size(); i++){
            const char c = text[i];
            if(c == '<'){ depth++; if(depth == 1) continue; }
            if(c == '>'){ depth--; if(depth == 0){ flush(); break; } }
            if(depth == 1 && c == ','){ flush(); continue; }
            if(depth == 1) tok += c; else if(depth > 1) tok += 'x';   // nested argument: make the token non-numeric
        }
        return v;
    }
    static std::vector<int64_t> integers_after(const std::string& text, const char* name){
        const size_t at = text.find(name);
        if(at == std::string::npos) return {};
        const size_t lt = text.find('<', at);
        return lt == std::string::npos ? std::vector<int64_t>{} : template_integers(text, lt);
    }

    bool memory_list(){
        // p is just after `memory`; expect [] = { ints }
        skip_ws(); if(p < end && *p == '['){ while(p < end && *p != ']') p++; p++; }
        skip_ws(); if(p >= end || *p != '=') return true;            // some other use of the word
        p++; skip_ws();
        if(p >= end || *p != '{') return fail("memory[]: expected '{'");
        p++;
        std::vector<unsigned char> bytes;
        for(;;){
            skip_ws();
            if(p >= end) return fail("memory[]: unterminated list");
            if(*p == '}'){ p++; break; }
            if(*p == ','){ p++; continue; }
            if(!std::isdigit((unsigned char)*p)) return fail("memory[]: expected a byte value");
            unsigned v = 0; while(p < end && std::isdigit((unsigned char)*p)){ v = v * 10 + (unsigned)(*p - '0'); p++; if(v > 255) return fail("memory[]: byte value out of range"); }
            bytes.push_back((unsigned char)v);
        }
        Ckpt::TensorEntry t;
        t.path = tensor_path(); t.bytes = (int64_t)bytes.size(); t.elem_size = pending_elem_size;
        if(t.elem_size != 4 && t.elem_size != 8) return fail("memory[]: unsupported element type");
        if(bytes.size() % (size_t)t.elem_size) return fail("memory[]: byte count is not a multiple of the element size");
        const size_t n = bytes.size() / (size_t)t.elem_size;
        t.data.resize(n);
        for(size_t i = 0; i < n; i++){
            if(t.elem_size == 4){ float f; std::memcpy(&f, bytes.data() + 4 * i, 4); t.data[i] = f; }
            else{ double d; std::memcpy(&d, bytes.data() + 8 * i, 8); t.data[i] = (float)d; }
        }
        out.tensors.push_back(std::move(t));
        pending_elem_size = 4;
        return true;
    }
    // `using NAME = ... ;` (the body may contain a struct definition with its own braces and semicolons)
    bool using_alias(){
        skip_ws();
        const std::string name = ident();
        std::string text; int depth = 0;
        while(p < end){
            const char c = *p;
            if(c == '"' || c == '\''){ std::string v; const char* s = p; if(!skip_string(&v)) return false; text.append(s, p); continue; }
            if(c == '{') depth++;
            if(c == '}') depth--;
            if(c == ';' && depth <= 0){ p++; break; }
            text += c; p++;
        }
        const std::string here = joined();
        if(name == "SHAPE" || name == "CONTAINER_SPEC"){
            if(out.tensors.empty() || out.tensors.back().shaped || out.tensors.back().path != tensor_path()) return true;
            auto& t = out.tensors.back();
            if(name == "SHAPE") t.dims = integers_after(text, "Shape");
            else{
                auto v = integers_after(text, "Specification");     // float, TI, ROWS, COLS, DYNAMIC, layout<...>
                if(v.size() >= 2) t.dims = {v[0], v[1]};
                auto a = integers_after(text, "RowMajorAlignment");
                if(!a.empty()) t.row_alignment = a.back();
            }
            if(t.dims.empty()) return fail("shape of '" + t.path + "' not understood");
            int64_t count = 1; for(auto d : t.dims) count *= d;
            const int64_t have = t.bytes / t.elem_size;
            if(have != count){
                // padded row pitch (matrix::layouts::RowMajorAlignment<TI, A> with A > 1): drop the padding
                if(t.dims.size() == 2 && t.row_alignment > 1){
                    const int64_t pitch = (t.dims[1] + t.row_alignment - 1) / t.row_alignment * t.row_alignment;
                    if(pitch * t.dims[0] == have){
                        std::vector<float> packed((size_t)count);
                        for(int64_t r = 0; r < t.dims[0]; r++) for(int64_t c = 0; c < t.dims[1]; c++) packed[(size_t)(r * t.dims[1] + c)] = t.data[(size_t)(r * pitch + c)];
                        t.data.swap(packed);
                    }
                    else return fail("'" + t.path + "': " + std::to_string((long long)have) + " values do not match its shape");
                }
                else return fail("'" + t.path + "': " + std::to_string((long long)have) + " values do not match its shape");
            }
            t.shaped = true;
        }
        else if(name == "CONFIG") out.config[here] = text;
        else if(name == "INPUT_SHAPE") out.input_shape[here] = integers_after(text, "Shape");
        return true;
    }

    bool run(){
        begin = p;
        while(true){
            skip_ws();
            if(p >= end) break;
            const char c = *p;
            if(c == '"' || c == '\''){ if(!skip_string(nullptr)) return false; continue; }
            if(c == '{'){ frames.push_back({0}); p++; continue; }
            if(c == '}'){
                if(frames.empty()) return fail("unbalanced '}'");
                for(int i = 0; i < frames.back().names; i++) path.pop_back();
                frames.pop_back(); p++; continue;
            }
            if(!ident_char(c) || std::isdigit((unsigned char)c)){ p++; continue; }
            const std::string word = ident();
            if(word == "namespace"){
                skip_ws();
                int names = 0;
                while(p < end && (ident_char(*p) || *p == ':')){
                    if(*p == ':'){ p++; continue; }
                    path.push_back(ident()); names++;
                }
                skip_ws();
                if(p >= end || *p != '{') return fail("namespace: expected '{'");
                p++;
                frames.push_back({names});
            }
            else if(word == "alignas"){
                skip_ws();
                if(p < end && *p == '('){ p++; skip_ws(); const std::string t = ident(); pending_elem_size = t == "double" ? 8 : t == "float" ? 4 : 0; while(p < end && *p != ')') p++; }
            }
            else if(word == "memory"){ if(!memory_list()) return false; }
            else if(word == "using"){ if(!using_alias()) return false; }
            else if(word == "char"){
                // meta strings: char name[] = "...";
                skip_ws(); const std::string name = ident(); skip_ws();
                if(name == "memory"){ if(!memory_list()) return false; continue; }   // `const unsigned char memory[] = {...}`
                if(p < end && *p == '['){ while(p < end && *p != ']') p++; p++; skip_ws(); if(p < end && *p == '='){ p++; skip_ws(); if(p < end && *p == '"'){ std::string v; if(!skip_string(&v)) return false; out.strings[joined() + "::" + name] = v; } } }
            }
        }
        if(!frames.empty()) return fail("unbalanced '{'");
        for(auto& t : out.tensors) if(!t.shaped) return fail("'" + t.path + "' has no shape declaration");
        return true;
    }
};

// ---- architecture recognition -----------------------------------------------------------------------------------------------------------------
struct Layer { std::string ns; enum Kind { DENSE, GRU, MLP, STANDARDIZE, SQUASH, UNKNOWN } kind = UNKNOWN; };

bool has(const Ckpt& c, const std::string& p){ return c.find(p) != nullptr; }
std::string activation_of(const Ckpt& c, const std::string& ns, int which = 0){
    auto it = c.config.find(ns);
    if(it == c.config.end()) return "";
    size_t at = 0; std::string name;
    for(int k = 0; k <= which; k++){
        at = it->second.find("ActivationFunction::", at);
        if(at == std::string::npos) return "";
        at += std::strlen("ActivationFunction::");
        size_t e = at; while(e < it->second.size() && (std::isalnum((unsigned char)it->second[e]) || it->second[e] == '_')) e++;
        name = it->second.substr(at, e - at);
    }
    return name;
}
Layer classify(const Ckpt& c, const std::string& ns){
    Layer l; l.ns = ns;
    auto cfg = c.config.find(ns);
    const std::string text = cfg == c.config.end() ? "" : cfg->second;
    if(has(c, ns + "::weights_input") && has(c, ns + "::weights_hidden")) l.kind = Layer::GRU;
    else if(has(c, ns + "::input_layer::weights") && has(c, ns + "::output_layer::weights")) l.kind = Layer::MLP;
    else if(has(c, ns + "::mean") && has(c, ns + "::precision")) l.kind = Layer::STANDARDIZE;
    else if(has(c, ns + "::weights") && has(c, ns + "::biases")) l.kind = Layer::DENSE;
    else if(text.find("sample_and_squash") != std::string::npos) l.kind = Layer::SQUASH;
    return l;
}
void append(std::vector<float>& blob, const Ckpt::TensorEntry* t){ blob.insert(blob.end(), t->data.begin(), t->data.end()); }

std::string build_policy(const Ckpt& c, const std::string& root, b200l2f_policy_desc& d, std::vector<float>& blob){
    std::vector<Layer> layers;
    for(int k = 0;; k++){
        const std::string ns = root + "::layer_" + std::to_string(k);
        Layer l = classify(c, ns);
        if(l.kind == Layer::UNKNOWN){
            bool any = false; for(auto& t : c.tensors) any = any || t.path.compare(0, ns.size() + 2, ns + "::") == 0;
            if(!any && c.config.find(ns) == c.config.end()) break;
            return "layer '" + ns + "' is of a kind the engine does not run";
        }
        layers.push_back(l);
    }
    if(layers.empty()){
        Layer l = classify(c, root);                       // a bare MLP saved without a sequential wrapper
        if(l.kind != Layer::MLP) return "no actor found under '" + root + "'";
        layers.push_back(l);
    }
    std::memset(&d, 0, sizeof(d));
    d.gemm = B200L2F_GEMM_TCGEN05_3XTF32;
    blob.clear();
    auto T = [&](const std::string& p){ return c.find(p); };
    if(layers.size() == 3 && layers[0].kind == Layer::DENSE && layers[1].kind == Layer::GRU && layers[2].kind == Layer::DENSE){
        const auto *w1 = T(layers[0].ns + "::weights"), *b1 = T(layers[0].ns + "::biases");
        const auto *wi = T(layers[1].ns + "::weights_input"), *bi = T(layers[1].ns + "::biases_input"), *wh = T(layers[1].ns + "::weights_hidden"), *bh = T(layers[1].ns + "::biases_hidden");
        const auto *h0 = T(layers[1].ns + "::initial_hidden_state");
        const auto *w2 = T(layers[2].ns + "::weights"), *b2 = T(layers[2].ns + "::biases");
        if(!w1 || !b1 || !wi || !wh || !w2 || !b2) return "Dense / GRU layer is missing weights or biases";
        if(!bi || !bh || !h0) return "GRU layer is missing biases_input / biases_hidden / initial_hidden_state";
        if(w1->dims.size() != 2 || wi->dims.size() != 2 || wh->dims.size() != 2 || w2->dims.size() != 2) return "weight tensors must be rank 2";
        const int64_t hid = w1->dims[0], in = w1->dims[1], out = w2->dims[0];
        if(wi->dims[0] != 3 * hid || wi->dims[1] != hid || wh->dims[0] != 3 * hid || wh->dims[1] != hid || w2->dims[1] != hid || (int64_t)b1->data.size() != hid ||
           (int64_t)bi->data.size() != 3 * hid || (int64_t)bh->data.size() != 3 * hid || (int64_t)h0->data.size() != hid || (int64_t)b2->data.size() != out)
            return "Dense/GRU/Dense shapes are inconsistent";
        if(activation_of(c, layers[0].ns) != "RELU" || activation_of(c, layers[2].ns) != "IDENTITY") return "engine runs Dense(ReLU) -> GRU -> Dense(identity) only";
        d.arch = B200L2F_POLICY_RAPTOR_GRU; d.input_dim = (int32_t)in; d.hidden_dim = (int32_t)hid; d.output_dim = (int32_t)out; d.head = B200L2F_HEAD_IDENTITY;
        auto is = c.input_shape.find(layers[1].ns);
        d.gru_sequence_length = is != c.input_shape.end() && !is->second.empty() ? (int32_t)is->second[0] : 0;   // SEQUENCE_LENGTH = first dim (gru/operations_generic.h:80)
        for(auto* t : {w1, b1, wi, bi, wh, bh, h0, w2, b2}) append(blob, t);
        return "";
    }
    size_t i = 0;
    const Ckpt::TensorEntry *mean = nullptr, *prec = nullptr;
    if(i < layers.size() && layers[i].kind == Layer::STANDARDIZE){ mean = T(layers[i].ns + "::mean"); prec = T(layers[i].ns + "::precision"); i++; }
    if(i >= layers.size() || layers[i].kind != Layer::MLP) return "unsupported layer sequence (expected Dense-GRU-Dense or [Standardize] MLP [SampleAndSquash])";
    const std::string m = layers[i].ns; i++;
    bool squash = false;
    if(i < layers.size() && layers[i].kind == Layer::SQUASH){ squash = true; i++; }
    if(i != layers.size()) return "unsupported layer after the MLP";
    if(has(c, m + "::hidden_layer_1::weights") || !has(c, m + "::hidden_layer_0::weights")) return "engine runs 3-layer MLPs (one hidden-to-hidden layer) only";
    const auto *w1 = T(m + "::input_layer::weights"), *b1 = T(m + "::input_layer::biases"), *w2 = T(m + "::hidden_layer_0::weights"), *b2 = T(m + "::hidden_layer_0::biases"),
               *w3 = T(m + "::output_layer::weights"), *b3 = T(m + "::output_layer::biases"), *ls = T(m + "::log_std");
    if(!w1 || !w2 || !w3) return "MLP layer without weights";
    if(!b1 || !b2 || !b3) return "MLP layer without biases";
    if(w1->dims.size() != 2 || w2->dims.size() != 2 || w3->dims.size() != 2) return "weight tensors must be rank 2";   // dims come from the file (.h5) / the export's SHAPE
    const int64_t hid = w1->dims[0], in = w1->dims[1], out = w3->dims[0];
    if(w2->dims[0] != hid || w2->dims[1] != hid || w3->dims[1] != hid) return "MLP shapes are inconsistent";
    if((int64_t)b1->data.size() != hid || (int64_t)b2->data.size() != hid || (int64_t)b3->data.size() != out || (ls && (int64_t)ls->data.size() != out))
        return "MLP bias / log_std sizes do not match the weights";
    if(mean && ((int64_t)mean->data.size() != in || (int64_t)prec->data.size() != in)) return "standardize statistics do not match the MLP input";
    const std::string hidden_act = activation_of(c, m, 0), out_act = activation_of(c, m, 1);
    if((!hidden_act.empty() && hidden_act != "RELU") || (!out_act.empty() && out_act != "IDENTITY")) return "engine runs ReLU hidden / identity output MLPs only";
    d.arch = B200L2F_POLICY_MLP; d.input_dim = (int32_t)in; d.hidden_dim = (int32_t)hid; d.output_dim = (int32_t)out; d.standardize = mean ? 1 : 0;
    d.head = squash ? B200L2F_HEAD_SQUASH_EVAL : ls ? B200L2F_HEAD_PPO_GAUSSIAN : B200L2F_HEAD_IDENTITY;
    if(mean){ append(blob, mean); append(blob, prec); }
    for(auto* t : {w1, b1, w2, b2, w3, b3}) append(blob, t);
    if(ls) append(blob, ls);
    return "";
}

int cfail(const std::string& m){ create_error() = m; return B200L2F_ERR_ARGUMENT; }

// ---- checkpoint.h5 -> the code export's namespace paths ---------------------------------------------------------------------------------------
// "/actor/layers/1/weights_input/parameters" -> "rl_tools::checkpoint::actor::layer_1::weights_input" (nn_models/sequential/persist.h:14-21 names the
// layer groups "layers/<k>", persist_code.h names the namespaces "layer_<k>"; nn/parameters/persist.h:10-13 stores a parameter's values as the
// dataset "parameters").  Other datasets of a parameter group ("gradient", optimizer moments of a Gradient-capability save) keep their name.
std::string h5_path_to_namespace(const std::string& path, bool drop_parameters){
    std::vector<std::string> parts;
    for(size_t i = 0; i < path.size();){
        while(i < path.size() && path[i] == '/') i++;
        size_t e = i; while(e < path.size() && path[e] != '/') e++;
        if(e > i) parts.push_back(path.substr(i, e - i));
        i = e;
    }
    if(drop_parameters && !parts.empty() && parts.back() == "parameters") parts.pop_back();
    std::string ns = "rl_tools::checkpoint";
    for(size_t i = 0; i < parts.size(); i++){
        const bool numbered = parts[i] == "layers" && i + 1 < parts.size() && !parts[i + 1].empty() && parts[i + 1].find_first_not_of("0123456789") == std::string::npos;
        if(numbered){ ns += "::layer_" + parts[i + 1]; i++; }
        else ns += "::" + parts[i];
    }
    return ns;
}
void from_h5(const H5Contents& f, Ckpt& c){
    for(auto& ds : f.datasets){
        Ckpt::TensorEntry t;
        t.path = h5_path_to_namespace(ds.path, true); t.dims = ds.dims; t.data = ds.data; t.elem_size = ds.elem_size; t.bytes = (int64_t)ds.data.size() * ds.elem_size; t.shaped = true;
        c.tensors.push_back(std::move(t));
    }
    // attributes: every string attribute is kept under <namespace>::<name>; the ones the code export states as template arguments are restated
    // in the form build_policy reads (dense/persist.h:17-18 "activation_function" / "type", sample_and_squash/persist.h:13, mlp/persist.h:16)
    std::vector<std::string> mlps;
    for(auto& a : f.attributes){
        const std::string ns = h5_path_to_namespace(a.object, true);
        c.strings[ns + "::" + a.name] = a.value;
        if(a.name == "activation_function") c.config[ns] = "ActivationFunction::" + a.value;
        else if(a.name == "type" && a.value == "sample_and_squash") c.config[ns] = "sample_and_squash";
        else if(a.name == "type" && a.value == "mlp") mlps.push_back(ns);
        else if(a.name == "checkpoint_name" && a.object == "/actor") c.strings["rl_tools::checkpoint::meta::name"] = a.value;
    }
    for(auto& m : mlps){                                              // hidden activation, then output activation (nn_models/mlp/network.h:15-51)
        auto hid = c.strings.find(m + "::input_layer::activation_function"), out = c.strings.find(m + "::output_layer::activation_function");
        if(hid != c.strings.end() && out != c.strings.end()) c.config[m] = "ActivationFunction::" + hid->second + " ActivationFunction::" + out->second;
    }
}

// No C++ exception may cross the C ABI (a corrupted file can still provoke std::length_error / bad_alloc inside the containers): every entry point
// that allocates or parses runs its body under guarded(), which turns an exception into an error code + message.
template <class F>
int guarded(const char* what, F&& body){
    try{ return body(); }
    catch(const std::exception& e){ return cfail(std::string(what) + ": " + e.what()); }
    catch(...){ return cfail(std::string(what) + ": unknown C++ exception"); }
}

}  // namespace

extern "C" {

int b200l2f_checkpoint_parse_h5(const void* bytes, size_t length, b200l2f_checkpoint** out){
    if(!bytes || !out) return cfail("checkpoint_parse_h5: null argument");
    *out = nullptr;
    return guarded("checkpoint_parse_h5", [&]() -> int {
        H5Contents f; std::string err;
        if(!h5_read((const unsigned char*)bytes, length, f, err)) return cfail("checkpoint_parse_h5: " + err);
        std::unique_ptr<b200l2f_checkpoint> c(new b200l2f_checkpoint());
        from_h5(f, *c);
        if(c->tensors.empty()) return cfail("checkpoint_parse_h5: the file holds no numeric datasets");
        *out = c.release();
        return B200L2F_OK;
    });
}
int b200l2f_checkpoint_parse(const char* text, size_t length, b200l2f_checkpoint** out){
    if(!text || !out) return cfail("checkpoint_parse: null argument");
    if(h5_has_signature((const unsigned char*)text, length)) return b200l2f_checkpoint_parse_h5(text, length, out);   // checkpoint.h5 handed to the generic entry
    *out = nullptr;
    return guarded("checkpoint_parse", [&]() -> int {
        std::unique_ptr<b200l2f_checkpoint> c(new b200l2f_checkpoint());
        Scanner s{text, text + length, *c, {}, {}, 4, text};
        if(!s.run()) return cfail("checkpoint_parse: " + c->err);
        if(c->tensors.empty()) return cfail("checkpoint_parse: no `memory[]` tensors found (not an rl-tools code export?)");
        *out = c.release();
        return B200L2F_OK;
    });
}
int b200l2f_checkpoint_free(b200l2f_checkpoint* c){ delete c; return B200L2F_OK; }
int b200l2f_checkpoint_tensor_count(const b200l2f_checkpoint* c){ return c ? (int)c->tensors.size() : 0; }
int b200l2f_checkpoint_tensor(const b200l2f_checkpoint* c, int index, const char** path, int32_t* rank, const int64_t** dims, const float** data){
    if(!c || index < 0 || index >= (int)c->tensors.size()) return cfail("checkpoint_tensor: index out of range");
    const auto& t = c->tensors[(size_t)index];
    if(path) *path = t.path.c_str();
    if(rank) *rank = (int32_t)t.dims.size();
    if(dims) *dims = t.dims.data();
    if(data) *data = t.data.data();
    return B200L2F_OK;
}
const char* b200l2f_checkpoint_string(const b200l2f_checkpoint* c, const char* path){
    if(!c || !path) return nullptr;
    auto it = c->strings.find(path);
    return it == c->strings.end() ? nullptr : it->second.c_str();
}
int b200l2f_checkpoint_string_count(const b200l2f_checkpoint* c){ return c ? (int)c->strings.size() : 0; }
int b200l2f_checkpoint_string_at(const b200l2f_checkpoint* c, int index, const char** path, const char** value){
    if(!c || index < 0 || index >= (int)c->strings.size()) return cfail("checkpoint_string_at: index out of range");
    auto it = c->strings.begin(); std::advance(it, index);
    if(path) *path = it->first.c_str();
    if(value) *value = it->second.c_str();
    return B200L2F_OK;
}
int b200l2f_checkpoint_policy(const b200l2f_checkpoint* c, const char* root, b200l2f_policy_desc* desc, float* blob, size_t capacity, size_t* n_floats){
    if(!c || !desc) return cfail("checkpoint_policy: null argument");
    return guarded("checkpoint_policy", [&]() -> int {
        std::vector<float> b; b200l2f_policy_desc d;
        const std::string err = build_policy(*c, root && *root ? root : "rl_tools::checkpoint::actor", d, b);
        if(!err.empty()){ create_error() = "checkpoint_policy: " + err; return B200L2F_ERR_UNSUPPORTED; }
        *desc = d;
        if(n_floats) *n_floats = b.size();
        if(blob){
            if(capacity < b.size()) return cfail("checkpoint_policy: blob capacity " + std::to_string(capacity) + " < " + std::to_string(b.size()) + " floats");
            std::memcpy(blob, b.data(), sizeof(float) * b.size());
        }
        return B200L2F_OK;
    });
}

}  // extern "C"
