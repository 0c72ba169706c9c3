// TEST INFRASTRUCTURE: stands in for the reference's include/misc/Exception.h (which needs the OptiX SDK headers) when its
// source is compiled on the CPU by oracle/build_ref.sh.
#pragma once
#include <stdexcept>
#include <string>
namespace psdr { struct Exception : std::runtime_error { using std::runtime_error::runtime_error; }; }
#define PSDR_ASSERT(cond) do { if (!(cond)) throw psdr::Exception(#cond); } while (0)
#define PSDR_ASSERT_MSG(cond, msg) do { if (!(cond)) throw psdr::Exception(msg); } while (0)
