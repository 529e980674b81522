// C driver around the reference's UNMODIFIED priority-queue MEX
// (hlc/optimizer/graph_search/priority_queue/priority_queue_interface_mex.cpp),
// compiled from where it lies under /root/reference via -DPQ_REF_SOURCE=...
// against oracle/mex_stub.  Output: oracle/_ref/libpq_ref.so (git-ignored).
// TEST INFRASTRUCTURE: tests/test_oracle_pq.py compares the oracle's heap
// restatement against this, push for push and pop for pop.
#include PQ_REF_SOURCE

#include <cstdint>

static MexFunction &mexfn() {
    static MexFunction f;
    return f;
}

extern "C" {

// PriorityQueue.m:19-21  this.iObj = mex(NEW)
int64_t pq_ref_new(void) {
    std::vector<matlab::data::Array> in(1), out(1);
    in[0].values = {0.0};  // NEW
    mexfn()(matlab::mex::ArgumentList(out), matlab::mex::ArgumentList(in));
    return static_cast<int64_t>(out[0].values[0]);
}

// PriorityQueue.m:27-29  mex(PUSH, iObj, ids, prio_vals)
void pq_ref_push(int64_t obj, const double *ids, const double *vals, int64_t n) {
    std::vector<matlab::data::Array> in(4), out(0);
    in[0].values = {2.0};  // PUSH
    in[1].values = {static_cast<double>(obj)};
    in[2].values.assign(ids, ids + n);
    in[3].values.assign(vals, vals + n);
    mexfn()(matlab::mex::ArgumentList(out), matlab::mex::ArgumentList(in));
}

// PriorityQueue.m:31-33  [id, val] = mex(POP, iObj); returns -1 when empty
int64_t pq_ref_pop(int64_t obj, double *val) {
    std::vector<matlab::data::Array> in(2), out(2);
    in[0].values = {3.0};  // POP
    in[1].values = {static_cast<double>(obj)};
    mexfn()(matlab::mex::ArgumentList(out), matlab::mex::ArgumentList(in));
    if (val) *val = out[1].values[0];
    return static_cast<int64_t>(out[0].values[0]);
}

// PriorityQueue.m:39-41
int64_t pq_ref_size(int64_t obj) {
    std::vector<matlab::data::Array> in(2), out(1);
    in[0].values = {5.0};  // SIZE
    in[1].values = {static_cast<double>(obj)};
    mexfn()(matlab::mex::ArgumentList(out), matlab::mex::ArgumentList(in));
    return static_cast<int64_t>(out[0].values[0]);
}
}
