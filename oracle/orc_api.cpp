// orc_api.cpp — allocation and field access of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
#include <cstring>

#include "orc_math.h"
#include "orc_types.h"

extern "C" {

OrcData* orc_make_data(const b2mjModel* m) {
  OrcData* d = new OrcData();
  auto allocd = [&](int n) { d->bufs.emplace_back((size_t)(n > 0 ? n : 1), 0.0); return d->bufs.back().data(); };
  auto alloci = [&](int n) { d->ibufs.emplace_back((size_t)(n > 0 ? n : 1), 0); return d->ibufs.back().data(); };
  struct Sel {
    static double* pick(double*, double* a, int*) { return a; }
    static int* pick(int*, double*, int* b) { return b; }
  };
#define X(e, n, t, c)                                              \
  {                                                                \
    double* pd = nullptr;                                          \
    int* pi = nullptr;                                             \
    if (sizeof(t) == sizeof(double)) pd = allocd(c);               \
    else pi = alloci(c);                                           \
    d->n = Sel::pick((t*)nullptr, pd, pi);                         \
  }
  ORC_FIELDS(X)
#undef X
  d->qH = allocd(m->nM);
  d->qHDiagInv = allocd(m->nv);
  d->contact_H = allocd(36 * m->nconmax);
  d->cb_control = nullptr;
  d->cb_passive = nullptr;
  d->cb_user = nullptr;
  d->n_control_calls = d->n_passive_calls = 0;
  orc_reset_data(m, d);
  return d;
}

void orc_free_data(OrcData* d) { delete d; }

void orc_set_callbacks(OrcData* d, orc_callback control, orc_callback passive, void* user) {
  d->cb_control = control;
  d->cb_passive = passive;
  d->cb_user = user;
}

int orc_callback_counts(const OrcData* d, int* ncontrol, int* npassive) {
  if (ncontrol) *ncontrol = d->n_control_calls;
  if (npassive) *npassive = d->n_passive_calls;
  return 0;
}

static int field_info(const b2mjModel* m, const OrcData* d, int field, void** ptr, int* elem) {
#define X(e, n, t, c)          \
  if (field == e) {            \
    *ptr = (void*)d->n;        \
    *elem = (int)sizeof(t);    \
    return (c);                \
  }
  ORC_FIELDS(X)
#undef X
  return -1;
}

int orc_get(const b2mjModel* m, const OrcData* d, int field, void* dst, int max_elems) {
  void* p;
  int es;
  int n = field_info(m, d, field, &p, &es);
  if (n < 0) return -1;
  int k = n < max_elems ? n : max_elems;
  if (k > 0) std::memcpy(dst, p, (size_t)k * es);
  return n;
}

int orc_set(const b2mjModel* m, OrcData* d, int field, const void* src, int nelems) {
  void* p;
  int es;
  int n = field_info(m, d, field, &p, &es);
  if (n < 0 || nelems > n) return -1;
  if (nelems > 0) std::memcpy(p, src, (size_t)nelems * es);
  return n;
}

void* orc_field_ptr(OrcData* d, int field) {
#define X(e, n, t, c) \
  if (field == e) return (void*)d->n;
  const b2mjModel* m = nullptr;
  (void)m;
  ORC_FIELDS(X)
#undef X
  return nullptr;
}

}  // extern "C"
