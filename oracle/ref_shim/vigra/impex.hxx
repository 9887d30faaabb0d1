// ORACLE — TEST INFRASTRUCTURE ONLY.  <vigra/impex.hxx> is included by sift.cpp:6 but nothing from it is
// used on the path (image import lives in main.cpp, which is not compiled).
#ifndef REF_SHIM_VIGRA_IMPEX_HXX
#define REF_SHIM_VIGRA_IMPEX_HXX
#endif
