// tests/stubs/napi_emul.cc -- see napi_emul.h.  TEST INFRASTRUCTURE, not product code.
#include "napi_emul.h"

#include <cstring>

using emul::Value;

static emul::Env g_env;
static napi_module *g_module = nullptr;
struct napi_callback_info__ {
    Value *self;
    std::vector<Value *> args;
    void *data;
};
static Value *V(napi_value v) { return reinterpret_cast<Value *>(v); }
static napi_value N(Value *v) { return reinterpret_cast<napi_value>(v); }
static Value *mk(napi_valuetype t) { Value *v = new Value(); v->type = t; return v; }
static size_t elem_size(napi_typedarray_type t) {
    switch (t) {
        case napi_int8_array: case napi_uint8_array: case napi_uint8_clamped_array: return 1;
        case napi_int16_array: case napi_uint16_array: return 2;
        case napi_int32_array: case napi_uint32_array: case napi_float32_array: return 4;
        default: return 8;
    }
}

namespace emul {
napi_env env() { return reinterpret_cast<napi_env>(&g_env); }
Value *num(double v) { Value *x = mk(napi_number); x->num = v; return x; }
Value *boolean(bool v) { Value *x = mk(napi_boolean); x->b = v; return x; }
Value *str(const char *s) { Value *x = mk(napi_string); x->str = s; return x; }
Value *null() { return mk(napi_null); }
Value *undefined() { return mk(napi_undefined); }
Value *object() { return mk(napi_object); }
Value *array(const std::vector<Value *> &e) { Value *x = mk(napi_object); x->is_array = true; x->elems = e; return x; }
static Value *typed(napi_typedarray_type t, const void *data, size_t n) {
    Value *x = mk(napi_object);
    x->is_typed = true; x->ta_type = t; x->ta_len = n;
    x->bytes = new std::vector<unsigned char>(n * elem_size(t));
    if (n) std::memcpy(x->bytes->data(), data, n * elem_size(t));
    return x;
}
Value *f64(const std::vector<double> &v) { return typed(napi_float64_array, v.data(), v.size()); }
Value *i32(const std::vector<int32_t> &v) { return typed(napi_int32_array, v.data(), v.size()); }
Value *u8(const std::vector<uint8_t> &v) { return typed(napi_uint8_array, v.data(), v.size()); }
Value *get(Value *obj, const char *key) {
    auto it = obj->props.find(key);
    return it == obj->props.end() ? nullptr : it->second;
}
std::vector<double> f64_of(Value *v) {
    std::vector<double> out(v->ta_len);
    if (v->ta_len) std::memcpy(out.data(), v->bytes->data() + v->ta_off, v->ta_len * 8);
    return out;
}
std::vector<int32_t> i32_of(Value *v) {
    std::vector<int32_t> out(v->ta_len);
    if (v->ta_len) std::memcpy(out.data(), v->bytes->data() + v->ta_off, v->ta_len * 4);
    return out;
}
static Value *invoke(napi_callback cb, void *data, Value *self, const std::vector<Value *> &args) {
    napi_callback_info__ info{self, args, data};
    napi_value r = cb(env(), &info);
    return g_env.pending ? nullptr : (r ? V(r) : undefined());
}
Value *construct(Value *ctor, const std::vector<Value *> &args) {
    Value *self = object();
    for (const napi_property_descriptor &m : ctor->methods)
        if (!(m.attributes & napi_static)) {
            Value *f = mk(napi_function);
            f->cb = m.method; f->cb_data = m.data;
            self->props[m.utf8name] = f;
        }
    return invoke(ctor->cb, ctor->cb_data, self, args) ? self : nullptr;
}
Value *call(Value *obj, const char *method, const std::vector<Value *> &args) {
    Value *f = get(obj, method);
    if (!f || f->type != napi_function) { g_env.pending = true; g_env.message = std::string("no such method: ") + method; return nullptr; }
    return invoke(f->cb, f->cb_data, obj, args);
}
bool exception_pending(std::string *message) {
    const bool p = g_env.pending;
    if (message) *message = g_env.message;
    g_env.pending = false; g_env.message.clear();
    return p;
}
Value *load_module() {
    if (!g_module) return nullptr;
    Value *exports = object();
    napi_value r = g_module->nm_register_func(env(), N(exports));
    return r ? V(r) : exports;
}
}  // namespace emul

extern "C" {
void napi_module_register(napi_module *mod) { g_module = mod; }
napi_status napi_get_undefined(napi_env, napi_value *r) { *r = N(emul::undefined()); return napi_ok; }
napi_status napi_get_null(napi_env, napi_value *r) { *r = N(emul::null()); return napi_ok; }
napi_status napi_get_boolean(napi_env, bool v, napi_value *r) { *r = N(emul::boolean(v)); return napi_ok; }
napi_status napi_typeof(napi_env, napi_value v, napi_valuetype *r) { *r = V(v)->type; return napi_ok; }
napi_status napi_coerce_to_bool(napi_env, napi_value v, napi_value *r) {
    Value *x = V(v);
    bool b = false;
    switch (x->type) {
        case napi_boolean: b = x->b; break;
        case napi_number: b = x->num != 0 && x->num == x->num; break;
        case napi_string: b = !x->str.empty(); break;
        case napi_undefined: case napi_null: b = false; break;
        default: b = true;
    }
    *r = N(emul::boolean(b));
    return napi_ok;
}
napi_status napi_get_value_bool(napi_env, napi_value v, bool *r) { if (V(v)->type != napi_boolean) return napi_boolean_expected; *r = V(v)->b; return napi_ok; }
napi_status napi_get_value_double(napi_env, napi_value v, double *r) { if (V(v)->type != napi_number) return napi_number_expected; *r = V(v)->num; return napi_ok; }
napi_status napi_get_value_string_utf8(napi_env, napi_value v, char *buf, size_t bufsize, size_t *result) {
    if (V(v)->type != napi_string) return napi_string_expected;
    const std::string &s = V(v)->str;
    if (!buf) { if (result) *result = s.size(); return napi_ok; }
    const size_t n = bufsize == 0 ? 0 : (s.size() < bufsize - 1 ? s.size() : bufsize - 1);
    std::memcpy(buf, s.data(), n);
    if (bufsize) buf[n] = 0;
    if (result) *result = n;
    return napi_ok;
}
napi_status napi_create_double(napi_env, double v, napi_value *r) { *r = N(emul::num(v)); return napi_ok; }
napi_status napi_create_int32(napi_env, int32_t v, napi_value *r) { *r = N(emul::num(v)); return napi_ok; }
napi_status napi_create_string_utf8(napi_env, const char *s, size_t len, napi_value *r) {
    Value *x = mk(napi_string);
    x->str = len == NAPI_AUTO_LENGTH ? std::string(s) : std::string(s, len);
    *r = N(x);
    return napi_ok;
}
napi_status napi_create_object(napi_env, napi_value *r) { *r = N(emul::object()); return napi_ok; }
napi_status napi_create_array_with_length(napi_env, size_t n, napi_value *r) {
    Value *x = emul::array({});
    x->elems.assign(n, nullptr);
    *r = N(x);
    return napi_ok;
}
napi_status napi_get_array_length(napi_env, napi_value v, uint32_t *r) { if (!V(v)->is_array) return napi_array_expected; *r = (uint32_t)V(v)->elems.size(); return napi_ok; }
napi_status napi_get_element(napi_env, napi_value o, uint32_t i, napi_value *r) {
    Value *x = V(o);
    *r = N(i < x->elems.size() && x->elems[i] ? x->elems[i] : emul::undefined());
    return napi_ok;
}
napi_status napi_set_element(napi_env, napi_value o, uint32_t i, napi_value v) {
    Value *x = V(o);
    if (i >= x->elems.size()) x->elems.resize(i + 1, nullptr);
    x->elems[i] = V(v);
    return napi_ok;
}
napi_status napi_set_named_property(napi_env, napi_value o, const char *k, napi_value v) { if (!v) return napi_invalid_arg; V(o)->props[k] = V(v); return napi_ok; }
napi_status napi_get_named_property(napi_env, napi_value o, const char *k, napi_value *r) {
    Value *p = emul::get(V(o), k);
    *r = N(p ? p : emul::undefined());
    return napi_ok;
}
napi_status napi_has_named_property(napi_env, napi_value o, const char *k, bool *r) { *r = emul::get(V(o), k) != nullptr; return napi_ok; }
napi_status napi_is_typedarray(napi_env, napi_value v, bool *r) { *r = V(v)->is_typed; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type *type, size_t *length, void **data, napi_value *ab, size_t *off) {
    Value *x = V(v);
    if (!x->is_typed) return napi_invalid_arg;
    if (type) *type = x->ta_type;
    if (length) *length = x->ta_len;
    if (data) *data = x->ta_len ? x->bytes->data() + x->ta_off : nullptr;
    if (ab) *ab = nullptr;
    if (off) *off = x->ta_off;
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env, size_t n, void **data, napi_value *r) {
    Value *x = mk(napi_object);
    x->is_buffer = true;
    x->bytes = new std::vector<unsigned char>(n);
    if (data) *data = x->bytes->data();
    *r = N(x);
    return napi_ok;
}
napi_status napi_create_typedarray(napi_env, napi_typedarray_type t, size_t len, napi_value ab, size_t off, napi_value *r) {
    Value *b = V(ab);
    if (!b->is_buffer || off + len * elem_size(t) > b->bytes->size()) return napi_invalid_arg;
    Value *x = mk(napi_object);
    x->is_typed = true; x->ta_type = t; x->ta_len = len; x->bytes = b->bytes; x->ta_off = off;
    *r = N(x);
    return napi_ok;
}
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t *argc, napi_value *argv, napi_value *self, void **data) {
    if (argc) {
        const size_t want = *argc;
        for (size_t i = 0; i < want && argv; i++) argv[i] = N(i < info->args.size() ? info->args[i] : emul::undefined());
        *argc = info->args.size();
    }
    if (self) *self = N(info->self);
    if (data) *data = info->data;
    return napi_ok;
}
napi_status napi_define_class(napi_env, const char *, size_t, napi_callback ctor, void *data, size_t n, const napi_property_descriptor *props, napi_value *r) {
    Value *c = mk(napi_function);
    c->cb = ctor; c->cb_data = data;
    c->methods.assign(props, props + n);
    for (size_t i = 0; i < n; i++)
        if (props[i].attributes & napi_static) {
            Value *f = mk(napi_function);
            f->cb = props[i].method; f->cb_data = props[i].data;
            c->props[props[i].utf8name] = f;
        }
    *r = N(c);
    return napi_ok;
}
napi_status napi_wrap(napi_env, napi_value o, void *native, napi_finalize, void *, napi_ref *) { V(o)->wrapped = native; return napi_ok; }
napi_status napi_unwrap(napi_env, napi_value o, void **r) { if (!V(o)->wrapped) return napi_invalid_arg; *r = V(o)->wrapped; return napi_ok; }
napi_status napi_create_reference(napi_env, napi_value v, uint32_t, napi_ref *r) { *r = reinterpret_cast<napi_ref>(v); return napi_ok; }
napi_status napi_create_function(napi_env, const char *, size_t, napi_callback cb, void *data, napi_value *r) {
    Value *f = mk(napi_function);
    f->cb = cb; f->cb_data = data;
    *r = N(f);
    return napi_ok;
}
static napi_status throw_(const char *msg) { g_env.pending = true; g_env.message = msg ? msg : ""; return napi_ok; }
napi_status napi_throw_error(napi_env, const char *, const char *msg) { return throw_(msg); }
napi_status napi_throw_type_error(napi_env, const char *, const char *msg) { return throw_(msg); }
napi_status napi_throw_range_error(napi_env, const char *, const char *msg) { return throw_(msg); }
}
