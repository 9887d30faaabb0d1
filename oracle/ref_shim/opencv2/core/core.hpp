// ORACLE — TEST INFRASTRUCTURE ONLY.  sift.cpp:9 includes this header and uses nothing from it.
