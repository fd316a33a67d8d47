// raptor_b200/csrc/pybind_l2f.cpp -- the `l2f` / `foundation_policy` Python modules of the README (R/README.md:19-24,40-105) as a pybind11 extension
// over the C ABI of include/b200_l2f.h: the compiled twin of raptor_b200/l2f.py + raptor_b200/foundation_policy.py (same names, argument orders and
// binding rules), i.e. the shim the reference ships as the pip wheels `l2f` / `foundation-policy` (not part of the reference tree).
//
//     import raptor_b200._l2f_pybind as l2f
//     vector = l2f.vector8                     # any N: l2f.vector(N)
//     policy = l2f.foundation_policy.Raptor()
//
// Host-only C++ (g++, pybind11 headers); links libb200l2f.so.  Objects keep the README's shapes: VectorEnvironment owns the engine handle,
// VectorParameters / VectorRng are tokens bound to it on first use, every VectorState maps to a state slot of the handle (or is a detached host
// snapshot after copy.copy, the README's `ui_state`).  Arrays: numpy float32, C-contiguous, host memory.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <cstdio>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200_l2f.h"

namespace py = pybind11;
using farray = py::array_t<float, py::array::c_style | py::array::forcecast>;

namespace {

constexpr int MAX_SLOTS = 8;

struct Engine {                              // one handle = N environments on one GPU
    b200l2f_handle* h = nullptr;
    int n = 0, obs_dim = 0, state_dim = 0, H = 0, slots = 0;
    Engine(int n_envs, int spec, int device){
        b200l2f_config c{(int32_t)sizeof(b200l2f_config), spec, n_envs, device, 0, MAX_SLOTS, 0, nullptr};
        if(b200l2f_create(&c, &h) != B200L2F_OK) throw std::runtime_error(std::string("b200l2f_create failed: ") + b200l2f_last_error(nullptr));
        n = n_envs; obs_dim = b200l2f_observation_dim(h); state_dim = b200l2f_state_dim(h); H = b200l2f_action_history_length(h);
    }
    ~Engine(){ if(h) b200l2f_destroy(h); }
    Engine(const Engine&) = delete;
    void check(int rc) const { if(rc != B200L2F_OK) throw std::runtime_error(std::string("b200l2f error ") + std::to_string(rc) + ": " + b200l2f_last_error(h)); }
    int alloc_slot(){ if(slots >= MAX_SLOTS) throw std::runtime_error("too many live VectorState objects for one VectorEnvironment"); return slots++; }
};
using EnginePtr = std::shared_ptr<Engine>;

struct Device { int ordinal = 0; };
struct UI { std::string ns; };
struct VectorEnvironment {
    EnginePtr e;
    int N_ENVIRONMENTS, OBSERVATION_DIM, ACTION_DIM = 4, EPISODE_STEP_LIMIT = 500;
    VectorEnvironment(int n, int spec, int device): e(std::make_shared<Engine>(n, spec, device)), N_ENVIRONMENTS(n), OBSERVATION_DIM(e->obs_dim) {}
};
struct VectorParameters { EnginePtr e; };
struct VectorRng { EnginePtr e; uint64_t seed = 0; };
struct VectorState {
    EnginePtr e; int slot = -1;
    std::vector<float> host; int host_rows = 0, host_cols = 0;      // detached snapshot (copy.copy(state))
    py::object host_array;                                          // numpy view of `host` handed to .states (kept alive with the object)
    int bind(const EnginePtr& env){
        if(!e){
            e = env; slot = env->alloc_slot();
            if(!host.empty()){ e->check(b200l2f_set_state(e->h, slot, host.data(), B200L2F_HOST)); host.clear(); host_array = py::none(); }
        }
        return slot;
    }
    farray numpy() const {
        if(!e){
            farray a({host_rows, host_cols});
            std::copy(host.begin(), host.end(), a.mutable_data());
            return a;
        }
        farray a({e->n, e->state_dim});
        e->check(b200l2f_get_state(e->h, slot, a.mutable_data(), B200L2F_HOST));
        return a;
    }
};
// host view of one environment's state row: s.position[0] += ... writes through to the snapshot (README render())
struct EnvState {
    py::object keep; float* row; int H;
    farray view(int offset, std::vector<py::ssize_t> shape) const {
        std::vector<py::ssize_t> strides(shape.size());
        py::ssize_t s = sizeof(float);
        for(int i = (int)shape.size() - 1; i >= 0; i--){ strides[(size_t)i] = s; s *= shape[(size_t)i]; }
        return farray(shape, strides, row + offset, keep);
    }
};

void bind_rng(const EnginePtr& env, VectorRng& rng){
    if(!rng.e){ rng.e = env; env->check(b200l2f_initialize_rng(env->h, rng.seed, 0)); }
}

std::string json_number(float v){ char b[64]; std::snprintf(b, sizeof(b), "%.9g", (double)v); return b; }
std::string json_list(const float* v, int n){ std::string s = "["; for(int i = 0; i < n; i++){ if(i) s += ", "; s += json_number(v[i]); } return s + "]"; }
std::string json_string(const std::string& v){ std::string s = "\""; for(char c : v){ if(c == '"' || c == '\\') s += '\\'; s += c; } return s + "\""; }

// one vectorN submodule: the README's `from l2f import vector8 as vector`
void define_vector_module(py::module_& m, int n){
    m.attr("N_ENVIRONMENTS") = n;
    m.def("VectorEnvironment", [n](int spec, int device){ return std::make_unique<VectorEnvironment>(n, spec, device); }, py::arg("spec") = (int)B200L2F_SPEC_DEFAULT, py::arg("device") = 0);
    m.def("VectorParameters", [](){ return std::make_unique<VectorParameters>(); });
    m.def("VectorRng", [](){ return std::make_unique<VectorRng>(); });
    m.def("VectorState", [](){ return std::make_unique<VectorState>(); });
    m.def("initialize_rng", [](Device&, VectorRng& rng, uint64_t seed){
        rng.seed = seed;
        if(rng.e) rng.e->check(b200l2f_initialize_rng(rng.e->h, seed, 0));
    });
    m.def("initialize_environment", [](Device&, VectorEnvironment& env){
        env.e->check(b200l2f_initialize_environment(env.e->h));
        env.e->check(b200l2f_initial_parameters(env.e->h));
    });
    m.def("sample_initial_parameters", [](Device&, VectorEnvironment& env, VectorParameters& p, VectorRng& rng){
        bind_rng(env.e, rng); p.e = env.e;
        env.e->check(b200l2f_sample_initial_parameters(env.e->h));
    });
    m.def("initial_parameters", [](Device&, VectorEnvironment& env, VectorParameters& p){ p.e = env.e; env.e->check(b200l2f_initial_parameters(env.e->h)); });
    m.def("sample_initial_state", [](Device&, VectorEnvironment& env, VectorParameters&, VectorState& s, VectorRng& rng){
        bind_rng(env.e, rng);
        env.e->check(b200l2f_sample_initial_state(env.e->h, s.bind(env.e)));
    });
    m.def("initial_state", [](Device&, VectorEnvironment& env, VectorParameters&, VectorState& s){ env.e->check(b200l2f_initial_state(env.e->h, s.bind(env.e))); });
    m.def("observe", [](Device&, VectorEnvironment& env, VectorParameters&, VectorState& s, py::array_t<float, py::array::c_style> observation, VectorRng& rng){
        bind_rng(env.e, rng);
        if(observation.ndim() != 2 || observation.shape(0) != env.e->n || observation.shape(1) < env.e->obs_dim) throw std::invalid_argument("observe: observation must be float32 [N_ENVIRONMENTS, >= OBSERVATION_DIM]");
        env.e->check(b200l2f_observe(env.e->h, s.bind(env.e), observation.mutable_data(), (int)observation.shape(1), B200L2F_HOST));   // in place, like the wheel
    });
    m.def("step", [](Device&, VectorEnvironment& env, VectorParameters&, VectorState& s, farray action, VectorState& next, VectorRng& rng){
        bind_rng(env.e, rng);
        if(action.ndim() != 2 || action.shape(0) != env.e->n || action.shape(1) != 4) throw std::invalid_argument("step: action must be [N_ENVIRONMENTS, 4]");
        std::vector<float> dts((size_t)env.e->n);
        env.e->check(b200l2f_step(env.e->h, s.bind(env.e), action.data(), next.bind(env.e), dts.data(), B200L2F_HOST));
        return std::vector<double>(dts.begin(), dts.end());        // list of dt (R/README.md:98,101)
    });
    // UI messages (R/README.md:72-77,86-88): same channels as L2F/ui.h:37-116, minimal payloads (no websocket in this repository)
    m.def("set_ui_message", [n](Device&, VectorEnvironment&, UI& ui){
        return "{\"namespace\": " + json_string(ui.ns) + ", \"channel\": \"setUI\", \"data\": {\"environments\": " + std::to_string(n) + "}}";
    });
    m.def("set_parameters_message", [](Device&, VectorEnvironment& env, VectorParameters&, UI& ui){
        std::vector<float> p((size_t)env.e->n * B200L2F_PARAMS_DIM);
        env.e->check(b200l2f_get_parameters(env.e->h, p.data(), B200L2F_HOST));
        std::string data = "[";
        for(int i = 0; i < env.e->n; i++){
            const float* r = p.data() + (size_t)i * B200L2F_PARAMS_DIM;
            if(i) data += ", ";
            data += "{\"parameters\": {\"dynamics\": {\"mass\": " + json_number(r[60]) + ", \"rotor_positions\": [";
            for(int k = 0; k < 4; k++){ if(k) data += ", "; data += json_list(r + 3 * k, 3); }
            data += "]}}}";
        }
        return "{\"namespace\": " + json_string(ui.ns) + ", \"channel\": \"setParameters\", \"data\": " + data + "]}";
    });
    m.def("set_state_action_message", [](Device&, VectorEnvironment&, VectorParameters&, UI& ui, VectorState& s, farray action){
        farray st = s.numpy();
        const int rows = (int)st.shape(0), cols = (int)st.shape(1);
        if(action.ndim() != 2 || action.shape(0) != rows || action.shape(1) != 4) throw std::invalid_argument("set_state_action_message: action must be [N_ENVIRONMENTS, 4]");
        std::string data = "[";
        for(int i = 0; i < rows; i++){
            const float* r = st.data() + (size_t)i * cols;
            if(i) data += ", ";
            data += "{\"state\": {\"position\": " + json_list(r, 3) + ", \"orientation\": " + json_list(r + 3, 4) + ", \"linear_velocity\": " + json_list(r + 7, 3) +
                    ", \"angular_velocity\": " + json_list(r + 10, 3) + ", \"rpm\": " + json_list(r + 26, 4) + "}, \"action\": " + json_list(action.data() + 4 * i, 4) + "}";
        }
        return "{\"namespace\": " + json_string(ui.ns) + ", \"channel\": \"setStateAction\", \"data\": " + data + "]}";
    });
}

// foundation_policy.Raptor (R/README.md:19-24,94-97): Mode<Evaluation>, hidden state auto-reset every 500 steps unless no_auto_reset
struct Raptor {
    int device; bool no_auto_reset; bool pending_reset = true;
    EnginePtr e;
    std::vector<float> blob; b200l2f_policy_desc desc{};
    Raptor(int device_, bool no_auto_reset_, const std::string& weights, const std::string& checkpoint): device(device_), no_auto_reset(no_auto_reset_){
        if(!checkpoint.empty()){                               // an rl-tools checkpoint.h code export
            std::ifstream f(checkpoint, std::ios::binary);
            if(!f) throw std::runtime_error("Raptor: cannot open " + checkpoint);
            std::stringstream ss; ss << f.rdbuf();
            const std::string text = ss.str();
            b200l2f_checkpoint* c = nullptr; size_t nf = 0;
            if(b200l2f_checkpoint_parse(text.data(), text.size(), &c) != B200L2F_OK) throw std::runtime_error(b200l2f_last_error(nullptr));
            int rc = b200l2f_checkpoint_policy(c, nullptr, &desc, nullptr, 0, &nf);
            if(rc == B200L2F_OK){ blob.resize(nf); rc = b200l2f_checkpoint_policy(c, nullptr, &desc, blob.data(), nf, &nf); }
            b200l2f_checkpoint_free(c);
            if(rc != B200L2F_OK) throw std::runtime_error(b200l2f_last_error(nullptr));
            if(desc.arch != B200L2F_POLICY_RAPTOR_GRU) throw std::runtime_error("Raptor: the checkpoint does not hold a Dense-GRU-Dense actor");
        }
        else{                                                  // the published checkpoint's weights shipped with the package (2084 floats)
            std::ifstream f(weights, std::ios::binary);
            if(!f) throw std::runtime_error("Raptor: cannot open " + weights);
            blob.resize(2084);
            f.read(reinterpret_cast<char*>(blob.data()), (std::streamsize)(blob.size() * sizeof(float)));
            if(f.gcount() != (std::streamsize)(blob.size() * sizeof(float))) throw std::runtime_error("Raptor: " + weights + " is not the 2084-parameter blob");
            desc = b200l2f_policy_desc{B200L2F_POLICY_RAPTOR_GRU, 22, 16, 4, 0, B200L2F_HEAD_IDENTITY, 500, B200L2F_GEMM_TCGEN05_3XTF32};
        }
    }
    void ensure(int n){
        if(!e || e->n != n){
            e = std::make_shared<Engine>(n, (int)B200L2F_SPEC_RAPTOR, device);
            e->check(b200l2f_policy_load(e->h, &desc, blob.data(), blob.size()));
            pending_reset = false;                             // policy_load resets
        }
    }
    void reset(){ if(!e) pending_reset = true; else e->check(b200l2f_policy_reset(e->h, nullptr, B200L2F_HOST)); }
    farray evaluate_step(farray obs){
        if(obs.ndim() != 2 || obs.shape(1) < desc.input_dim)
            throw std::invalid_argument("Raptor.evaluate_step expects [batch, 22] observations (position, rotation matrix, linear velocity, angular velocity, previous action)");
        ensure((int)obs.shape(0));
        if(pending_reset){ e->check(b200l2f_policy_reset(e->h, nullptr, B200L2F_HOST)); pending_reset = false; }
        farray act({(py::ssize_t)obs.shape(0), (py::ssize_t)4});
        e->check(b200l2f_policy_evaluate_step(e->h, obs.data(), (int)obs.shape(1), act.mutable_data(), no_auto_reset ? 1 : 0, B200L2F_HOST));
        return act;
    }
};

}  // namespace

PYBIND11_MODULE(_l2f_pybind, m){
    m.doc() = "README-compatible l2f / foundation_policy modules over the B200 rollout engine (pybind11 twin of raptor_b200.l2f / raptor_b200.foundation_policy)";
    py::class_<Device>(m, "Device").def(py::init<>()).def_readwrite("ordinal", &Device::ordinal);
    py::class_<UI>(m, "UI").def(py::init<>()).def_readwrite("ns", &UI::ns);
    py::class_<VectorEnvironment>(m, "_VectorEnvironment")
        .def_readonly("N_ENVIRONMENTS", &VectorEnvironment::N_ENVIRONMENTS).def_readonly("OBSERVATION_DIM", &VectorEnvironment::OBSERVATION_DIM)
        .def_readonly("ACTION_DIM", &VectorEnvironment::ACTION_DIM).def_readonly("EPISODE_STEP_LIMIT", &VectorEnvironment::EPISODE_STEP_LIMIT)
        .def_property_readonly("kernel_launches", [](VectorEnvironment& v){ return (long long)b200l2f_kernel_launches(v.e->h); });
    py::class_<VectorParameters>(m, "_VectorParameters");
    py::class_<VectorRng>(m, "_VectorRng");
    py::class_<EnvState>(m, "_EnvState")
        .def_property_readonly("position", [](EnvState& s){ return s.view(0, {3}); }).def_property_readonly("orientation", [](EnvState& s){ return s.view(3, {4}); })
        .def_property_readonly("linear_velocity", [](EnvState& s){ return s.view(7, {3}); }).def_property_readonly("angular_velocity", [](EnvState& s){ return s.view(10, {3}); })
        .def_property_readonly("last_action", [](EnvState& s){ return s.view(13, {4}); }).def_property_readonly("force", [](EnvState& s){ return s.view(20, {3}); })
        .def_property_readonly("torque", [](EnvState& s){ return s.view(23, {3}); }).def_property_readonly("rpm", [](EnvState& s){ return s.view(26, {4}); })
        .def_property_readonly("current_step", [](EnvState& s){ return (int)s.row[30]; })
        .def_property_readonly("action_history", [](EnvState& s){ return s.view(31, {s.H, 4}); });
    py::class_<VectorState>(m, "_VectorState")
        .def("assign", [](VectorState& self, VectorState& other){
            if(!other.e) throw std::runtime_error("assign: source state has never been used with an environment");
            self.bind(other.e);
            self.e->check(b200l2f_copy_state(self.e->h, self.slot, other.slot));
        })
        .def("numpy", &VectorState::numpy)
        .def_property_readonly("states", [](py::object self_obj){
            VectorState& self = self_obj.cast<VectorState&>();
            if(self.host.empty()){                              // first access: take the snapshot the views write through to
                farray a = self.numpy();
                self.host.assign(a.data(), a.data() + a.size());
                self.host_rows = (int)a.shape(0); self.host_cols = (int)a.shape(1);
            }
            const int H = (self.host_cols - 44) / 4;
            py::list out;
            for(int i = 0; i < self.host_rows; i++) out.append(EnvState{self_obj, self.host.data() + (size_t)i * self.host_cols, H});
            return out;
        })
        .def("__copy__", [](VectorState& self){
            auto c = std::make_unique<VectorState>();
            farray a = self.numpy();
            c->host.assign(a.data(), a.data() + a.size());
            c->host_rows = (int)a.shape(0); c->host_cols = (int)a.shape(1);
            return c;
        });
    m.def("vector", [m](int n) mutable {                        // l2f.vector(N): the vectorN module for any N (created on first use)
        const std::string name = "vector" + std::to_string(n);
        if(py::hasattr(m, name.c_str())) return m.attr(name.c_str()).cast<py::module_>();
        py::module_ sub = m.def_submodule(name.c_str());
        define_vector_module(sub, n);
        return sub;
    });
    for(int n : {1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 4096, 16384, 65536}){   // the wheel's fixed set; any other N through vector(N)
        py::module_ sub = m.def_submodule(("vector" + std::to_string(n)).c_str());
        define_vector_module(sub, n);
    }
    py::module_ fp = m.def_submodule("foundation_policy");
    py::class_<Raptor>(fp, "Raptor")
        .def(py::init([](int device, bool no_auto_reset, const std::string& checkpoint){
            py::object here = py::module_::import("os").attr("path").attr("dirname")(py::module_::import("raptor_b200").attr("__file__"));
            const std::string weights = py::str(here).cast<std::string>() + "/data/raptor_policy_2084.f32";
            return std::make_unique<Raptor>(device, no_auto_reset, weights, checkpoint);
        }), py::arg("device") = 0, py::arg("no_auto_reset") = false, py::arg("checkpoint") = "")
        .def("reset", &Raptor::reset)
        .def("evaluate_step", &Raptor::evaluate_step);
    for(auto kv : {std::pair<const char*, int>{"SPEC_DEFAULT", B200L2F_SPEC_DEFAULT}, {"SPEC_DEFAULT_DR", B200L2F_SPEC_DEFAULT_DR}, {"SPEC_RAPTOR", B200L2F_SPEC_RAPTOR},
                   {"SPEC_TEACHER", B200L2F_SPEC_TEACHER}, {"SPEC_RAPTOR_DR", B200L2F_SPEC_RAPTOR_DR}, {"SPEC_TEACHER_DR", B200L2F_SPEC_TEACHER_DR}})
        m.attr(kv.first) = kv.second;
}
