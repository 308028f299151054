// thread-local error string behind mc_last_error()
#pragma once
#include <string>
namespace mc {
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);
}  // namespace mc
