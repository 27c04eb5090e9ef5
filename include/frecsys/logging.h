// Minimal stand-in for the glog LOG(INFO)/LOG(ERROR) stream macros the reference uses
// (glog 0.6.0 is not a dependency of this build): one line per statement on stderr with a
// glog-style prefix, so log scrapers written for the reference keep working.
#pragma once
#include <chrono>
#include <cstdio>
#include <ctime>
#include <sstream>
#include <string>

namespace frecsys {
namespace logging {
class Line {
public:
  Line(char sev, const char* file, int line) {
    using namespace std::chrono;
    const auto now = system_clock::now();
    const std::time_t t = system_clock::to_time_t(now);
    std::tm tmv;
    localtime_r(&t, &tmv);
    const long us = (long)(duration_cast<microseconds>(now.time_since_epoch()).count() % 1000000);
    const char* base = file;
    for (const char* p = file; *p; ++p)
      if (*p == '/') base = p + 1;
    char buf[96];
    std::snprintf(buf, sizeof buf, "%c%02d%02d %02d:%02d:%02d.%06ld %s:%d] ", sev, tmv.tm_mon + 1, tmv.tm_mday,
                  tmv.tm_hour, tmv.tm_min, tmv.tm_sec, us, base, line);
    ss_ << buf;
  }
  ~Line() {
    ss_ << "\n";
    std::fputs(ss_.str().c_str(), stderr);
  }
  std::ostringstream& stream() { return ss_; }
private:
  std::ostringstream ss_;
};
}  // namespace logging
}  // namespace frecsys

#ifndef LOG
#define FRECSYS_SEV_INFO 'I'
#define FRECSYS_SEV_WARNING 'W'
#define FRECSYS_SEV_ERROR 'E'
#define LOG(sev) ::frecsys::logging::Line(FRECSYS_SEV_##sev, __FILE__, __LINE__).stream()
#endif
