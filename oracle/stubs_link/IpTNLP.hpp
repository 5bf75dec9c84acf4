// Stand-in for Ipopt's IpTNLP.hpp for the link test (see IpJournalist.hpp here): include/sqphot/SQPTNLP.hpp only needs the name.
#ifndef ORACLE_STUB_LINK_IPTNLP_HPP
#define ORACLE_STUB_LINK_IPTNLP_HPP
#define ORACLE_STUB_IPTNLP_HPP
#include <IpJournalist.hpp>
namespace Ipopt {
class TNLP;
typedef int Index;
typedef double Number;
}
#endif
