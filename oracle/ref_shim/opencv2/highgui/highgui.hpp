// ORACLE — TEST INFRASTRUCTURE ONLY.  sift.cpp:10 includes this header and uses nothing from it.
