/* world/macrodefinitions.h -- linkage macros kept for source compatibility with callers of
 * mmorise/World (reference: src/world/macrodefinitions.h:66-74, 118-130).  Every public symbol
 * is a plain C symbol so that C99 programs (the reference's test/ctest.c) and the existing
 * language bindings relink unchanged. */
#ifndef WORLD_MACRODEFINITIONS_H_
#define WORLD_MACRODEFINITIONS_H_

#ifdef __cplusplus
#define WORLD_BEGIN_C_DECLS extern "C" {
#define WORLD_END_C_DECLS }
#else
#define WORLD_BEGIN_C_DECLS
#define WORLD_END_C_DECLS
#endif

#if defined(__GNUC__) && __GNUC__ >= 4
#define WORLD_API __attribute__((visibility("default")))
#else
#define WORLD_API
#endif

#endif /* WORLD_MACRODEFINITIONS_H_ */
