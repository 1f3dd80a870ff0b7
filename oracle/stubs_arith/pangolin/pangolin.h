// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the two Pangolin facilities MonoSLAM::Init uses
// (monoslam.cpp:1578-1850): pangolin::ParseVarsFile ("key = value;" lines, '#' / '%' comments) and
// pangolin::Var<T>(name, default) with implicit conversion to T.  No GUI.
#ifndef SL2_ORACLE_PANGOLIN_STUB
#define SL2_ORACLE_PANGOLIN_STUB
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>

namespace pangolin {

inline std::map<std::string, std::string> &var_store() {
  static std::map<std::string, std::string> s;
  return s;
}

inline void ParseVarsFile(const std::string &path) {
  std::ifstream f(path.c_str());
  std::string line;
  while (std::getline(f, line)) {
    const size_t c = line.find_first_of("#%");
    if (c != std::string::npos) line = line.substr(0, c);
    const size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string k = line.substr(0, eq), v = line.substr(eq + 1);
    const size_t semi = v.find(';');
    if (semi != std::string::npos) v = v.substr(0, semi);
    auto trim = [](std::string &s) {
      const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
      s = (a == std::string::npos) ? std::string() : s.substr(a, b - a + 1);
    };
    trim(k);
    trim(v);
    if (!k.empty()) var_store()[k] = v;
  }
}

template <class T>
inline T var_convert(const std::string &s) {
  std::istringstream is(s);
  T v = T();
  is >> v;
  return v;
}
template <>
inline std::string var_convert<std::string>(const std::string &s) { return s; }
template <>
inline bool var_convert<bool>(const std::string &s) { return s == "true" || s == "1"; }
template <>
inline int var_convert<int>(const std::string &s) { return (int)std::atof(s.c_str()); }

template <class T>
class Var {
 public:
  Var(const std::string &name, const T &def) : v_(def) {
    auto it = var_store().find(name);
    if (it != var_store().end()) v_ = var_convert<T>(it->second);
  }
  operator const T &() const { return v_; }
  const T &Get() const { return v_; }

 private:
  T v_;
};

}  // namespace pangolin
#endif
