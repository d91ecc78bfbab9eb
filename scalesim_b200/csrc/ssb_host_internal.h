// ssb_host_internal.h — what the host-only translation units of the library (scenario_host.cc) share with the engine
// (ssb_engine.cu) besides the public C ABI: the error strings behind ssb_last_error. Not exported.
#pragma once
#include <string>

struct ssb_engine;
// error of a call that has no engine (what ssb_last_error(NULL) returns; thread-local like ssb_create's)
__attribute__((visibility("hidden"))) void ssb_internal_set_error(const std::string& what);
__attribute__((visibility("hidden"))) void ssb_internal_set_engine_error(ssb_engine* e, const std::string& what);
