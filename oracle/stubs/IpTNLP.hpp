// Stand-in for Ipopt's IpTNLP.hpp (oracle/_ref build only; see IpJournalist.hpp stub).
#ifndef ORACLE_STUB_IPTNLP_HPP
#define ORACLE_STUB_IPTNLP_HPP
#include <IpJournalist.hpp>
#endif
