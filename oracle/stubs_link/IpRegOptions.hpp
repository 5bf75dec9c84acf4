// Stand-in for Ipopt's IpRegOptions.hpp (link tests): src/Algorithm.cpp registers its option table here and never reads it back.
#ifndef ORACLE_STUB_LINK_IPREGOPTIONS_HPP
#define ORACLE_STUB_LINK_IPREGOPTIONS_HPP
#include <string>
#include <IpJournalist.hpp>
namespace Ipopt {
class RegisteredOptions {
    typedef double Number_;
public:
    void SetRegisteringCategory(const std::string&) {}
    void AddNumberOption(const std::string&, const std::string&, Number_ = 0.0, const std::string& = "") {}
    void AddIntegerOption(const std::string&, const std::string&, int = 0, const std::string& = "") {}
    template <typename... T> void AddStringOption1(T...) {}
    template <typename... T> void AddStringOption2(T...) {}
    template <typename... T> void AddStringOption3(T...) {}
    template <typename... T> void AddStringOption4(T...) {}
    template <typename... T> void AddLowerBoundedNumberOption(T...) {}
    template <typename... T> void AddBoundedNumberOption(T...) {}
    template <typename... T> void AddLowerBoundedIntegerOption(T...) {}
    template <typename... T> void AddBoundedIntegerOption(T...) {}
};
}
#endif
