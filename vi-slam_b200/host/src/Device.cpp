// Device.cpp — process-wide libvislam_b200 context for the class mirrors (see device.hpp).
#include "vislam/device.hpp"

#include <cstdlib>
#include <sstream>

namespace vi {

Device::Device() : ctx_(nullptr), stream_(nullptr) {
    int ordinal = 0;
    if (const char* e = std::getenv("VISLAM_DEVICE")) ordinal = std::atoi(e);
    int rc = vsb_ctx_create(ordinal, &ctx_);
    if (rc != VSB_OK) {
        std::ostringstream os;
        os << "vislam_b200: cannot create a context on CUDA device " << ordinal << " (" << vsb_error_string(rc)
           << "); this path has no CPU fallback";
        throw DeviceError(os.str());
    }
    rc = vsb_stream_create(ctx_, &stream_);
    if (rc != VSB_OK) {
        std::string msg = std::string("vislam_b200: stream creation failed: ") + vsb_last_cuda_error(ctx_);
        vsb_ctx_destroy(ctx_);
        ctx_ = nullptr;
        throw DeviceError(msg);
    }
}

Device::~Device() {
    if (ctx_) {
        if (stream_) vsb_stream_destroy(ctx_, stream_);
        vsb_ctx_destroy(ctx_);
    }
}

Device& Device::get() {
    static Device* dev = new Device();   // never destroyed: buffers owned by static objects may outlive main()
    return *dev;
}

void Device::sync() const { check(vsb_stream_sync(ctx_, stream_), "stream synchronize"); }

void Device::check(int status, const char* what) const {
    if (status == VSB_OK) return;
    std::ostringstream os;
    os << "vislam_b200: " << what << " failed: " << vsb_error_string(status);
    if (status == VSB_ERR_CUDA) os << " (" << vsb_last_cuda_error(ctx_) << ")";
    throw DeviceError(os.str());
}

void* DevBuf::reserve(size_t bytes) {
    if (bytes <= cap_ && p_) return p_;
    Device& d = Device::get();
    if (p_) {
        d.sync();
        d.check(vsb_free(d.ctx(), p_), "vsb_free");
        p_ = nullptr;
        cap_ = 0;
    }
    const size_t want = bytes + bytes / 4 + 256;
    d.check(vsb_malloc(d.ctx(), want, &p_), "vsb_malloc");
    cap_ = want;
    return p_;
}

void DevBuf::reset() {
    if (!p_) return;
    try {
        Device& d = Device::get();
        d.sync();
        vsb_free(d.ctx(), p_);
    } catch (...) {
    }
    p_ = nullptr;
    cap_ = 0;
}

}  // namespace vi
