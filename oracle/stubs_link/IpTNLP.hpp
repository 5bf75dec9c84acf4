// Stand-in for Ipopt's IpTNLP.hpp for the link tests (see IpJournalist.hpp here): the abstract NLP interface with the eight
// callbacks src/SQPTNLP.cpp calls, in Ipopt's published signatures.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_LINK_IPTNLP_HPP
#define ORACLE_STUB_LINK_IPTNLP_HPP
#define ORACLE_STUB_IPTNLP_HPP
#include <IpJournalist.hpp>
namespace Ipopt {
typedef int Index;
typedef double Number;
class TNLP {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
    virtual bool get_nlp_info(Index& n, Index& m, Index& nnz_jac_g, Index& nnz_h_lag, IndexStyleEnum& index_style) = 0;
    virtual bool get_bounds_info(Index n, Number* x_l, Number* x_u, Index m, Number* g_l, Number* g_u) = 0;
    virtual bool get_starting_point(Index n, bool init_x, Number* x, bool init_z, Number* z_L, Number* z_U, Index m, bool init_lambda,
                                    Number* lambda) = 0;
    virtual bool eval_f(Index n, const Number* x, bool new_x, Number& obj_value) = 0;
    virtual bool eval_grad_f(Index n, const Number* x, bool new_x, Number* grad_f) = 0;
    virtual bool eval_g(Index n, const Number* x, bool new_x, Index m, Number* g) = 0;
    virtual bool eval_jac_g(Index n, const Number* x, bool new_x, Index m, Index nele_jac, Index* iRow, Index* jCol, Number* values) = 0;
    virtual bool eval_h(Index n, const Number* x, bool new_x, Number obj_factor, Index m, const Number* lambda, bool new_lambda,
                        Index nele_hess, Index* iRow, Index* jCol, Number* values) = 0;
};
}
#endif
