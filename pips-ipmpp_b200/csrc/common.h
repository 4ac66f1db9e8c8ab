// Shared helpers of the C-ABI implementation: error capture (no exception crosses the C boundary).
#pragma once
#include <exception>
#include <stdexcept>
#include <string>

#include "../../include/pipsgpu.h"

namespace pipsgpu {
void set_error(const std::string& s);

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& s) : std::runtime_error(s), code(c) {}
};
}  // namespace pipsgpu

#define PIPSGPU_TRY try {
#define PIPSGPU_CATCH                                     \
    return PIPSGPU_OK;                                    \
    }                                                     \
    catch (const pipsgpu::Error& e) {                     \
        pipsgpu::set_error(e.what());                     \
        return e.code;                                    \
    }                                                     \
    catch (const std::exception& e) {                     \
        pipsgpu::set_error(e.what());                     \
        return PIPSGPU_ERR_INVALID;                       \
    }                                                     \
    catch (...) {                                         \
        pipsgpu::set_error("unknown exception");          \
        return PIPSGPU_ERR_INVALID;                       \
    }
