/* TEST INFRASTRUCTURE ONLY.  The reference generates git-revision.c with cmake's configure_file
 * (src/git-revision.c.in); our hand-written oracle recipe does not run cmake, so the five strings
 * that src/git-revision.h declares are provided here. */
const char GIT_BRANCH[] = "oracle-build";
const char GIT_REVISION_HASH[] = "unknown";
const char HOST_NAME[] = "oracle";
const char BUILD_DATE[] = __DATE__;
const char BUILD_TIME[] = __TIME__;
