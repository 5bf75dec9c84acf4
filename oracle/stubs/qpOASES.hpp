// Empty stand-in: the reference's L0 headers include <qpOASES.hpp> but use nothing from it.
// qpOASES 3.2.1 itself is NOT available in this container ("oracle").
#ifndef ORACLE_STUB_QPOASES_HPP
#define ORACLE_STUB_QPOASES_HPP
#endif
