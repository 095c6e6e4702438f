// device.hpp — the class mirrors' handle on libvislam_b200: one context + one stream per process (the
// reference is strictly single-threaded and synchronous), RAII device buffers, and status
// checking that fails LOUDLY: there is no CPU fallback behind these classes.
#ifndef VISLAM_DEVICE_HPP_
#define VISLAM_DEVICE_HPP_

#include <cstddef>
#include <stdexcept>
#include <string>

#include "vislam_b200.h"

namespace vi {

class DeviceError : public std::runtime_error {
public:
    explicit DeviceError(const std::string& what) : std::runtime_error(what) {}
};

class Device {
public:
    // The process-wide device (ordinal from $VISLAM_DEVICE, default 0).  Throws DeviceError when no CUDA device
    // or no libvislam_b200 context can be had.
    static Device& get();
    vsb_ctx_t* ctx() const { return ctx_; }
    void* stream() const { return stream_; }
    void sync() const;
    void check(int status, const char* what) const;
    long long launches() const { return vsb_launch_count(ctx_); }

private:
    Device();
    ~Device();
    Device(const Device&);
    Device& operator=(const Device&);
    vsb_ctx_t* ctx_;
    void* stream_;
};

// Device allocation that grows on demand and never shrinks (per-frame sizes are stable).
class DevBuf {
public:
    DevBuf() : p_(nullptr), cap_(0) {}
    ~DevBuf() { reset(); }
    void* reserve(size_t bytes);          // contents are NOT preserved when the buffer grows
    void reset();
    void* get() const { return p_; }
    template <typename T> T* as() const { return static_cast<T*>(p_); }
    size_t capacity() const { return cap_; }

private:
    DevBuf(const DevBuf&);
    DevBuf& operator=(const DevBuf&);
    void* p_;
    size_t cap_;
};

}  // namespace vi
#endif
