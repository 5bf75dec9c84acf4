// Stand-in for Ipopt's IpOptionsList.hpp (link tests): src/Algorithm.cpp creates one and hands it out, nothing else.
#ifndef ORACLE_STUB_LINK_IPOPTIONSLIST_HPP
#define ORACLE_STUB_LINK_IPOPTIONSLIST_HPP
#include <IpJournalist.hpp>
namespace Ipopt {
class OptionsList {};
}
#endif
