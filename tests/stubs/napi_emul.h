// tests/stubs/napi_emul.h -- TEST INFRASTRUCTURE: a minimal in-process emulation of the N-API functions declared in
// tests/stubs/node_api.h, enough to EXECUTE binding/jslp_addon.cc without Node.js (tests/cpp/addon_emul_test.cc).
// Values are heap objects that live until process exit; there is no garbage collector, no event loop, no JS.
#pragma once
#include <node_api.h>

#include <map>
#include <string>
#include <vector>

namespace emul {

struct Value {
    napi_valuetype type = napi_undefined;
    double num = 0;
    bool b = false;
    std::string str;
    std::map<std::string, Value *> props;
    std::vector<Value *> elems;          // arrays
    bool is_array = false;
    // typed array / arraybuffer
    bool is_typed = false, is_buffer = false;
    napi_typedarray_type ta_type = napi_uint8_array;
    size_t ta_len = 0;
    std::vector<unsigned char> *bytes = nullptr;  // shared with the arraybuffer
    size_t ta_off = 0;
    // function / class
    napi_callback cb = nullptr;
    void *cb_data = nullptr;
    std::vector<napi_property_descriptor> methods;  // class: instance + static methods
    // wrap
    void *wrapped = nullptr;
};

struct Env {
    bool pending = false;
    std::string message;
};

napi_env env();
Value *num(double v);
Value *boolean(bool v);
Value *str(const char *s);
Value *null();
Value *undefined();
Value *object();
Value *array(const std::vector<Value *> &e);
Value *f64(const std::vector<double> &v);
Value *i32(const std::vector<int32_t> &v);
Value *u8(const std::vector<uint8_t> &v);
Value *get(Value *obj, const char *key);  // nullptr when absent
std::vector<double> f64_of(Value *v);
std::vector<int32_t> i32_of(Value *v);
// `new ctor(args...)`; nullptr (and a pending exception) on failure
Value *construct(Value *ctor, const std::vector<Value *> &args);
// obj.method(args...) / ctor.staticMethod(args...)
Value *call(Value *obj, const char *method, const std::vector<Value *> &args);
bool exception_pending(std::string *message);  // reads and clears
Value *load_module();                          // runs the registered module's init, returns `exports`

}  // namespace emul
