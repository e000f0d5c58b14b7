// Stub of google-glog for the oracle build (test infrastructure only, see oracle/README.md).
// The reference only uses LOG/DLOG/LOG_IF/DLOG_IF streams; FATAL aborts like glog does.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace ne_oracle_stub {
struct NullStream { template <class T> NullStream& operator<<(const T&) { return *this; } };
struct FatalStream {
	std::ostringstream ss;
	template <class T> FatalStream& operator<<(const T& v) { ss << v; return *this; }
	~FatalStream() { std::cerr << "[reference LOG(FATAL)] " << ss.str() << std::endl; std::abort(); }
};
}
#define NE_STUB_LOG_INFO ne_oracle_stub::NullStream()
#define NE_STUB_LOG_WARNING ne_oracle_stub::NullStream()
#define NE_STUB_LOG_ERROR ne_oracle_stub::NullStream()
#define NE_STUB_LOG_FATAL ne_oracle_stub::FatalStream()
#define LOG(sev) NE_STUB_LOG_##sev
#define DLOG(sev) ne_oracle_stub::NullStream()
#define LOG_IF(sev, cond) if (cond) NE_STUB_LOG_##sev
#define DLOG_IF(sev, cond) if (false) ne_oracle_stub::NullStream()
namespace google { inline void InitGoogleLogging(const char*) {} }
