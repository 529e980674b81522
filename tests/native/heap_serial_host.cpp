// Host build of p-dmpc_b200/csrc/pdmpc_heap_serial.h for tests/test_heap_serial.py:
// the product's single-thread heap routines behind a tiny C interface, so that
// they can be compared push for push / pop for pop with the reference's own MEX.
#include <cstdint>
#include <vector>

#include "../../p-dmpc_b200/csrc/pdmpc_heap_serial.h"

namespace {
struct Ent {
    double f;
    long long id;
};
struct Q {
    std::vector<Ent> a;
    int len = 0;
};
}  // namespace

extern "C" {
void *hs_new() { return new Q(); }
void hs_free(void *q) { delete static_cast<Q *>(q); }
void hs_push(void *qv, long long id, double f) {
    Q *q = static_cast<Q *>(qv);
    if ((int)q->a.size() <= q->len) q->a.resize(q->len * 2 + 16);
    pdmpc::heap_push_serial<Ent>(q->a.data(), q->len, Ent{f, id});
    ++q->len;
}
long long hs_pop(void *qv, double *f) {
    Q *q = static_cast<Q *>(qv);
    if (q->len == 0) return -1;
    Ent e = pdmpc::heap_pop_serial<Ent>(q->a.data(), q->len);
    --q->len;
    if (f) *f = e.f;
    return e.id;
}
int hs_size(void *qv) { return static_cast<Q *>(qv)->len; }
// array state, for the stronger check "same array after every operation"
long long hs_id_at(void *qv, int i) { return static_cast<Q *>(qv)->a[i].id; }
}
