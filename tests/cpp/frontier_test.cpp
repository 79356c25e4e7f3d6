// Replays the ordering scenarios of the reference's min-heap.test.ts:150-292 and the all-gather record round trip
// against the PRODUCT's frontier structures (jslpsolver_b200/csrc/jslp_frontier.h).  Built and run by
// tests/test_host_cpu.py with g++; prints FRONTIER OK.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "jslp_b200.h"
#include "jslp_frontier.h"

using namespace jslp_bnb;

#define CHECK(c) do { if (!(c)) { std::printf("FAILED line %d: %s\n", __LINE__, #c); std::exit(1); } } while (0)

static Branch *push(Frontier &f, double ev) {
    std::unique_ptr<Branch> b(new Branch{ev, {}, NodeEval()});
    Branch *raw = b.get();
    f.push(std::move(b));
    return raw;
}
static double pop_eval(Frontier &f) { return f.pop_entry().b->relaxedEvaluation; }
static Branch *pop_ptr(Frontier &f) { Frontier::Entry e = f.pop_entry(); return e.b.release(); }

int main() {
    {   // returns elements in sorted order (min first)
        Frontier f;
        for (double v : {30.0, 10.0, 20.0, 5.0, 25.0}) push(f, v);
        for (double v : {5.0, 10.0, 20.0, 25.0, 30.0}) CHECK(pop_eval(f) == v);
        CHECK(f.empty());
    }
    {   // LIFO tie-breaking: the most recently pushed of equal evaluations comes out first
        Frontier f;
        Branch *first = push(f, 5), *second = push(f, 5), *third = push(f, 5);
        Branch *p;
        CHECK((p = pop_ptr(f)) == third); delete p;
        CHECK((p = pop_ptr(f)) == second); delete p;
        CHECK((p = pop_ptr(f)) == first); delete p;
    }
    {   // combines min-heap and LIFO correctly
        Frontier f;
        Branch *a = push(f, 10), *b = push(f, 5), *c = push(f, 5), *d = push(f, 3);
        Branch *p;
        CHECK((p = pop_ptr(f)) == d); delete p;
        CHECK((p = pop_ptr(f)) == c); delete p;
        CHECK((p = pop_ptr(f)) == b); delete p;
        CHECK((p = pop_ptr(f)) == a); delete p;
    }
    {   // maintains heap property after interleaved operations
        Frontier f;
        push(f, 50); push(f, 30);
        CHECK(pop_eval(f) == 30);
        push(f, 20); push(f, 40);
        CHECK(pop_eval(f) == 20);
        push(f, 10);
        CHECK(pop_eval(f) == 10);
        CHECK(pop_eval(f) == 40);
        CHECK(pop_eval(f) == 50);
        CHECK(f.empty());
    }
    {   // negative and fractional evaluations
        Frontier f;
        for (double v : {-10.0, -30.0, -20.0, 0.0}) push(f, v);
        for (double v : {-30.0, -20.0, -10.0, 0.0}) CHECK(pop_eval(f) == v);
        for (double v : {1.5, 1.1, 1.3}) push(f, v);
        for (double v : {1.1, 1.3, 1.5}) CHECK(pop_eval(f) == v);
    }
    {   // re-inserting an entry with its ORIGINAL sequence number (what a speculative round does) keeps the order
        Frontier f;
        push(f, 7); push(f, 7); push(f, 7);
        Frontier::Entry top = f.pop_entry();       // seq 2
        Frontier::Entry next = f.pop_entry();      // seq 1
        CHECK(top.seq == 2 && next.seq == 1);
        f.push_entry(std::move(next));
        f.push_entry(std::move(top));
        CHECK(f.pop_entry().seq == 2);
        CHECK(f.pop_entry().seq == 1);
        CHECK(f.pop_entry().seq == 0);
    }
    {   // cuts survive push and pop
        Frontier f;
        std::unique_ptr<Branch> b(new Branch{1.0, {}, NodeEval()});
        b->cuts.push_back(jslp_cut{0, 3, 2.0});
        b->cuts.push_back(jslp_cut{1, 4, 7.0});
        f.push(std::move(b));
        Frontier::Entry e = f.pop_entry();
        CHECK(e.b->cuts.size() == 2 && e.b->cuts[0].type == 0 && e.b->cuts[0].var_index == 3 && e.b->cuts[1].value == 7.0);
    }
    {   // the 128-byte all-gather record is lossless, including -inf evaluations and optional-objective entries
        NodeEval e;
        e.valid = true; e.feasible = 1; e.bounded = 0; e.optimal = 0; e.is_integral = 1; e.branch_var = 17; e.pivots = 12345;
        e.evaluation = -INFINITY; e.branch_value = 2.5000000000000004;
        for (int o = 0; o < 7; o++) e.opt0[o] = 0.1 * (o + 1);
        double w[WIRE_DOUBLES];
        to_wire(e, w);
        NodeEval r;
        from_wire(r, w);
        CHECK(r.valid && r.feasible == 1 && r.bounded == 0 && r.optimal == 0 && r.is_integral == 1);
        CHECK(r.branch_var == 17 && r.pivots == 12345 && std::isinf(r.evaluation) && r.evaluation < 0);
        CHECK(r.branch_value == 2.5000000000000004);
        for (int o = 0; o < 7; o++) CHECK(r.opt0[o] == 0.1 * (o + 1));
        CHECK(sizeof(w) == 128);
    }
    std::printf("FRONTIER OK\n");
    return 0;
}
