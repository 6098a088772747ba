// errors.h -- error plumbing shared by the translation units of libnirrt_b200.so
#pragma once
#include <string>
// stores the message returned by nirrt_last_error() for the calling thread; returns `code`
int nirrt_set_error(int code, const std::string &msg);
