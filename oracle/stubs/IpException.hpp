// Stand-in for Ipopt's IpException.hpp: just enough for include/sqphot/QPsolverInterface.hpp:26-32
// (DECLARE_STD_EXCEPTION / THROW_EXCEPTION).  Compile-check infrastructure only.
#ifndef ORACLE_STUB_IPEXCEPTION_HPP
#define ORACLE_STUB_IPEXCEPTION_HPP
#include <IpJournalist.hpp>
#include <string>
namespace Ipopt {
class IpoptException {
public:
    IpoptException(std::string msg, std::string file, int line, std::string type = "IpoptException")
        : msg_(msg), file_(file), line_(line), type_(type) {}
    virtual ~IpoptException() {}
    const std::string& Message() const { return msg_; }
private:
    std::string msg_, file_;
    int line_;
    std::string type_;
};
}  // namespace Ipopt
#define THROW_EXCEPTION(__except_type, __msg) throw __except_type((__msg), (__FILE__), (__LINE__));
#define DECLARE_STD_EXCEPTION(__except_type)                                                        \
    class __except_type : public Ipopt::IpoptException {                                            \
    public:                                                                                         \
        __except_type(std::string msg, std::string fname, int line) : Ipopt::IpoptException(msg, fname, line, #__except_type) {} \
    }
#endif
