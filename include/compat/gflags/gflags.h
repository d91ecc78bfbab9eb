// Stand-in for <gflags/gflags.h>: DEFINE_bool / DEFINE_string register FLAGS_*
// globals; ParseCommandLineFlags consumes "--name", "--name=value", "--noname"
// (and "-name" forms) and, with remove_flags, compacts argv to the positionals.
#pragma once
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
namespace gflags {
struct FlagRegistry {
  std::map<std::string, bool*> bools;
  std::map<std::string, std::string*> strings;
  std::string usage;
  static FlagRegistry& get() { static FlagRegistry r; return r; }
};
struct BoolRegisterer { BoolRegisterer(const char* n, bool* p) { FlagRegistry::get().bools[n] = p; } };
struct StringRegisterer { StringRegisterer(const char* n, std::string* p) { FlagRegistry::get().strings[n] = p; } };
inline void SetUsageMessage(const std::string& u) { FlagRegistry::get().usage = u; }
inline unsigned ParseCommandLineFlags(int* argc, char*** argv, bool remove_flags) {
  FlagRegistry& r = FlagRegistry::get();
  int out = 1;
  for (int i = 1; i < *argc; ++i) {
    char* a = (*argv)[i];
    if (a[0] != '-' || a[1] == '\0' || (a[1] >= '0' && a[1] <= '9')) { (*argv)[out++] = a; continue; }
    std::string s(a + (a[1] == '-' ? 2 : 1));
    if (s == "help") { std::cerr << r.usage << std::endl; std::exit(0); }
    std::string name = s, val;
    bool has_val = false;
    std::size_t eq = s.find('=');
    if (eq != std::string::npos) { name = s.substr(0, eq); val = s.substr(eq + 1); has_val = true; }
    if (val.size() >= 2 && val.front() == '"' && val.back() == '"') val = val.substr(1, val.size() - 2);
    if (r.bools.count(name)) {
      *r.bools[name] = !has_val || (val != "false" && val != "0" && val != "no");
    } else if (name.compare(0, 2, "no") == 0 && r.bools.count(name.substr(2))) {
      *r.bools[name.substr(2)] = false;
    } else if (r.strings.count(name)) {
      if (!has_val && i + 1 < *argc) { val = (*argv)[++i]; }
      *r.strings[name] = val;
    } else {
      std::cerr << "ERROR: unknown command line flag '" << name << "'" << std::endl;
      std::exit(1);
    }
  }
  if (remove_flags) { *argc = out; (*argv)[out] = nullptr; }
  return static_cast<unsigned>(out);
}
}  // namespace gflags
#define DEFINE_bool(name, val, txt) \
  bool FLAGS_##name = val; \
  static ::gflags::BoolRegisterer scalesim_compat_flagreg_##name(#name, &FLAGS_##name)
#define DEFINE_string(name, val, txt) \
  std::string FLAGS_##name = val; \
  static ::gflags::StringRegisterer scalesim_compat_flagreg_##name(#name, &FLAGS_##name)
#define DECLARE_bool(name) extern bool FLAGS_##name
#define DECLARE_string(name) extern std::string FLAGS_##name
