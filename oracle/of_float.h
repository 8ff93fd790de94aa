/* TEST INFRASTRUCTURE ONLY: replaces the body of the reference's src/of.h (switched off with
 * -DOF_OF_H) for the float cross-check build of the unmodified reference sources. */
#ifndef ORACLE_OF_FLOAT_H
#define ORACLE_OF_FLOAT_H
typedef float ofpix_t;
#endif
