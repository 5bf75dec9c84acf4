// Stand-in for Ipopt's IpJournalist.hpp for the LINK test of the reference's own QPhandler.cpp / qpOASESInterface.cpp /
// QOREInterface.cpp (oracle/Makefile target _ref/qphandler_hs071): the stub of oracle/stubs plus the journal-file calls those
// translation units make.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_STUB_LINK_IPJOURNALIST_HPP
#define ORACLE_STUB_LINK_IPJOURNALIST_HPP
#define ORACLE_STUB_IPJOURNALIST_HPP  // keeps oracle/stubs/IpJournalist.hpp out
#include <cstdarg>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <string>
namespace Ipopt {
enum EJournalLevel { J_INSUPPRESSIBLE = -1, J_NONE = 0, J_ERROR, J_STRONGWARNING, J_SUMMARY, J_WARNING,
                     J_ITERSUMMARY, J_DETAILED, J_MOREDETAILED, J_VECTOR, J_MOREVECTOR, J_MATRIX,
                     J_MOREMATRIX, J_ALL, J_LAST_LEVEL };
enum EJournalCategory { J_DBG = 0, J_STATISTICS, J_MAIN, J_INITIALIZATION, J_BARRIER_UPDATE,
                        J_SOLVE_PD_SYSTEM, J_FRAC_TO_BOUND, J_LINEAR_ALGEBRA, J_LINE_SEARCH,
                        J_HESSIAN_APPROXIMATION, J_SOLUTION, J_DOCUMENTATION, J_NLP, J_TIMING_STATISTICS,
                        J_USER_APPLICATION, J_USER1, J_USER2, J_LAST_CATEGORY };
template <class T> class SmartPtr {
public:
    SmartPtr() : p_(nullptr) {}
    SmartPtr(T* p) : p_(p) {}
    SmartPtr(std::nullptr_t) : p_(nullptr) {}
    T* operator->() const { return p_; }
    T& operator*() const { return *p_; }
    T* get() const { return p_; }
private:
    T* p_;
};
template <class T> inline bool IsValid(const SmartPtr<T>& p) { return p.get() != nullptr; }
template <class T> inline bool IsNull(const SmartPtr<T>& p) { return p.get() == nullptr; }
template <class T> inline T* GetRawPtr(const SmartPtr<T>& p) { return p.get(); }
class Journal {
public:
    void SetPrintLevel(EJournalCategory, EJournalLevel) {}
    void SetAllPrintLevels(EJournalLevel) {}
};
class Journalist {
public:
    // silent unless ORACLE_STUB_JOURNAL_STDOUT is set: the reference prints an iteration log through this call
    void Printf(EJournalLevel, EJournalCategory, const char* fmt, ...) const {
        static const bool loud = getenv("ORACLE_STUB_JOURNAL_STDOUT") != nullptr;
        if (!loud) return;
        va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap);
    }
    SmartPtr<Journal> AddFileJournal(const std::string&, const std::string&, EJournalLevel = J_WARNING) { return SmartPtr<Journal>(&journal_); }
    SmartPtr<Journal> GetJournal(const std::string&) { return SmartPtr<Journal>(&journal_); }
    void DeleteAllJournals() {}
    void FlushBuffer() const {}
private:
    Journal journal_;
};
}  // namespace Ipopt
#endif
