// tests/cpp/addon_emul_test.cc -- executes binding/jslp_addon.cc (the N-API shim) on the GPU through the in-process N-API
// emulation of tests/stubs/napi_emul.cc: every addon method is called the way gpu-tableau.ts calls it, on the README
// Berlin Airlift LP, an integer model and an edited tableau, and its results are compared with the known answers
// (README.md:73).  Built and run by tests/test_gpu_addon.py on the GPU box (no Node.js there).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "napi_emul.h"

using namespace emul;

static int fails = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); fails++; } \
    } while (0)
static void no_exception(const char *where) {
    std::string msg;
    if (exception_pending(&msg)) { std::printf("FAIL unexpected exception at %s: %s\n", where, msg.c_str()); fails++; }
}
static Value *opts(std::initializer_list<std::pair<const char *, Value *>> kv) {
    Value *o = object();
    for (auto &p : kv) o->props[p.first] = p.second;
    return o;
}

int main() {
    Value *exports = load_module();
    CHECK(exports != nullptr);
    Value *Tab = get(exports, "Tab");
    CHECK(Tab && Tab->type == napi_function);
    Value *abi = call(exports, "abiVersion", {});
    CHECK(abi && abi->num == 2);

    // ---- README Berlin Airlift: W = 3, H = 4
    Value *t = construct(Tab, {num(3), num(4), num(8), num(1e-8)});
    no_exception("new Tab");
    CHECK(t != nullptr);
    if (!t) return 1;
    call(t, "upload", {f64({0, 20000, 30000, 44, 1, 1, 512, 8, 16, 300000, 5000, 9000}), i32({-1, 0, 1, 2}), i32({-1, 3, 4}),
                       null(), null(), null(), num(0)});
    no_exception("upload");
    call(t, "setOption", {num(3 /* JSLP_OPT_PIVOT_LOG_CAP */), num(64)});
    Value *s = call(t, "simplex", {boolean(true)});
    no_exception("simplex");
    CHECK(s && get(s, "feasible")->b && get(s, "bounded")->b && get(s, "evaluation")->num == -1080000);
    CHECK(s && get(s, "phase1Pivots")->num == 0 && get(s, "phase2Pivots")->num == 2);
    Value *d = call(t, "download", {opts({{"rhs", boolean(true)}, {"maps", boolean(true)}, {"matrix", boolean(true)}, {"cost", boolean(true)}})});
    no_exception("download");
    CHECK(d && get(d, "width")->num == 3 && get(d, "height")->num == 4);
    if (d) {
        const std::vector<double> rhs = f64_of(get(d, "rhs"));
        const std::vector<int32_t> vr = i32_of(get(d, "varIndexByRow")), vc = i32_of(get(d, "varIndexByCol"));
        CHECK(rhs.size() == 4 && rhs[0] == -1080000 && rhs[1] == 24 && rhs[2] == 20 && rhs[3] == 0);
        CHECK(vr == (std::vector<int32_t>{-1, 3, 4, 2}) && vc == (std::vector<int32_t>{-1, 0, 1}));
        CHECK(f64_of(get(d, "matrix")).size() == 12 && f64_of(get(d, "cost"))[0] == -1080000);
    }
    Value *log = call(t, "pivotLog", {});
    CHECK(log && i32_of(log).size() == 8 && i32_of(log)[0] == 2 && i32_of(log)[1] == 2 && i32_of(log)[4] == 1 && i32_of(log)[5] == 1);
    // save / cuts / restore: brit <= 20 through applyCuts, then back
    call(t, "save", {});
    Value *cut = opts({{"type", str("max")}, {"varIndex", num(3)}, {"value", num(20)}});
    Value *s2 = call(t, "applyCuts", {array({cut}), boolean(true)});
    no_exception("applyCuts");
    CHECK(s2 && get(s2, "feasible")->b && get(s2, "height")->num == 5 && get(s2, "evaluation")->num > -1080000);
    call(t, "restore", {});
    Value *i0 = call(t, "info", {});
    CHECK(i0 && get(i0, "height")->num == 4 && get(i0, "lastElementIndex")->num == 5);
    // dynamic modification: 44 planes -> 40 (the `plane` slack, index 0, is non-basic after the solve)
    call(t, "updateRhs", {num(0), num(-(40 - 44))});  // Constraint.setRightHandSide: difference = -(new - old) for an upper bound
    no_exception("updateRhs");
    Value *s3 = call(t, "simplex", {boolean(true)});
    CHECK(s3 && get(s3, "feasible")->b && get(s3, "evaluation")->num == -(20000.0 * 16 + 30000.0 * 24));  // brit 16, yank 24
    call(t, "destroy", {});

    // ---- integer model (tables / dressers with ints; the true optimum is 19200, not the 14400 README.md:147 prints)
    Value *m = construct(Tab, {num(3), num(3), num(16), num(1e-8)});
    no_exception("new Tab (MIP)");
    call(m, "upload", {f64({0, 1200, 1600, 300, 30, 20, 110, 5, 10}), i32({-1, 0, 1}), i32({-1, 2, 3}), null(), i32({2, 3}), null(), num(0)});
    no_exception("upload (MIP)");
    Value *r = call(m, "branchAndCut", {opts({{"tolerance", num(0)}, {"isMinimization", boolean(false)}, {"checkCycles", boolean(true)},
                                              {"keepSolutions", boolean(true)}})});
    no_exception("branchAndCut");
    CHECK(r && get(r, "feasible")->b && get(r, "isIntegral")->b && get(r, "evaluation")->num == -19200);
    CHECK(r && get(r, "iterations")->num >= 1 && get(r, "bestCuts")->is_array && get(r, "solutions")->is_array);
    Value *frac = call(m, "isIntegral", {});
    CHECK(frac && frac->b);
    Value *mf = call(m, "mostFractional", {});
    CHECK(mf && get(mf, "index")->type == napi_null);
    // the enhanced service through the same entry point
    call(m, "upload", {f64({0, 1200, 1600, 300, 30, 20, 110, 5, 10}), i32({-1, 0, 1}), i32({-1, 2, 3}), null(), i32({2, 3}), null(), num(0)});
    Value *r2 = call(m, "branchAndCut", {opts({{"isMinimization", boolean(false)}, {"nodeSelection", str("depth-first")}, {"branching", str("strong")}})});
    no_exception("branchAndCut (enhanced)");
    CHECK(r2 && get(r2, "evaluation")->num == -19200);

    // ---- misuse becomes a thrown Error, not a crash
    call(m, "upload", {f64({1, 2, 3}), i32({-1, 0, 1}), i32({-1, 2, 3}), null(), null(), null(), num(0)});
    std::string msg;
    CHECK(exception_pending(&msg) && msg.find("height * width") != std::string::npos);
    call(m, "pivot", {num(99), num(1)});
    CHECK(exception_pending(&msg) && msg.find("jslp_b200") != std::string::npos);
    Value *id = call(Tab, "uniqueId", {});
    if (exception_pending(&msg)) std::printf("note: uniqueId unavailable: %s\n", msg.c_str());  // no libnccl: allowed
    else CHECK(id && id->ta_len == 128);

    std::printf(fails ? "ADDON EMUL FAILED (%d)\n" : "ADDON EMUL OK\n", fails);
    return fails ? 1 : 0;
}
