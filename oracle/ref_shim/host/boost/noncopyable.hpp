#ifndef CLSIM_REF_SHIM_NONCOPYABLE_HPP
#define CLSIM_REF_SHIM_NONCOPYABLE_HPP
namespace boost { class noncopyable { protected: noncopyable() {} ~noncopyable() {} noncopyable(const noncopyable &) = delete; noncopyable &operator=(const noncopyable &) = delete; }; }
#endif
