// model_core.h — allocation / copy / reflection helpers for b2mjModel (library-internal).
#pragma once
#include <string>

#include "b2mj.h"

namespace b2mj {

// thread-local error string behind b2mj_last_error()
void set_error(const std::string& msg);

// allocate a zeroed model; arrays are allocated by model_alloc_arrays once the size fields are set
b2mjModel* model_new();
void model_alloc_arrays(b2mjModel* m);
b2mjModel* model_clone(const b2mjModel* src);
void model_default_option(b2mjOption* o);

// qpos0-dependent constants (mj_setConst semantics; reference re-runs it at callbacks.cpp:254,582)
int model_set_const(b2mjModel* m, std::string& err);

// static candidate geom-pair table (deterministic order: body-pair signature, then geom ids)
void model_build_collision_pairs(b2mjModel* m);

}  // namespace b2mj
