/* Force-included (-include) when compiling the UNMODIFIED reference sources in place.
 *
 * TEST INFRASTRUCTURE ONLY -- nothing under oracle/ is part of the product path.
 *
 * The reference writes `while(getline(fin, line)>0)` (inputReader/readLoader.cpp:86,
 * matePair/matePair.cpp:83).  Under C++11 a stream only converts to bool explicitly, so the
 * expression no longer compiles with g++ 13.  Rather than patching a copy of the sources, this
 * header supplies the one missing operator with the pre-C++11 meaning ("stream still good").
 */
#ifndef SAGE2_ORACLE_REF_COMPAT_H
#define SAGE2_ORACLE_REF_COMPAT_H
#ifdef __cplusplus
#include <istream>
inline bool operator>(std::istream &s, int) { return static_cast<bool>(s); }
#endif
#endif
