#pragma once
#include <stdarg.h>
#include <string>
namespace av1b {
extern thread_local std::string g_last_error;
void set_error(const char* fmt, ...);
}
