// ORACLE — TEST INFRASTRUCTURE ONLY.  Type names graphic/graphictool.h mentions; nothing is drawn.
#ifndef SL2_ORACLE_GL_STUB
#define SL2_ORACLE_GL_STUB
typedef unsigned int GLuint;
struct GLUquadricObj;
#endif
