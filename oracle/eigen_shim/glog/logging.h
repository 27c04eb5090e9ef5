// Stand-in for glog's LOG(severity) stream macros (test infrastructure only).
#pragma once
#include <iostream>
#include <sstream>
namespace google {
inline void InstallFailureSignalHandler() {}
struct ShimLogLine {
  std::ostringstream ss;
  bool quiet;
  explicit ShimLogLine(bool q) : quiet(q) {}
  ~ShimLogLine() { if (!quiet) std::cerr << ss.str() << "\n"; }
};
inline bool& ShimQuiet() { static bool q = false; return q; }
}  // namespace google
#define LOG(sev) ::google::ShimLogLine(::google::ShimQuiet()).ss
