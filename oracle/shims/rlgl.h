/* Fake "rlgl.h" — TEST INFRASTRUCTURE ONLY.  tau_hypersonic_3d_cuda.cu includes rlgl for its
 * slice-stack volume renderer; nothing from it is reached by oracle/ref_drivers/ref_hyp3d.cu.
 * Symbols are added here only as the compiler asks for them.  Contains no reference code. */
#ifndef TAU_FAKE_RLGL_H
#define TAU_FAKE_RLGL_H
#define RL_QUADS 7
#define RL_SRC_ALPHA 0x0302
#define RL_ONE_MINUS_SRC_ALPHA 0x0303
#define RL_ONE 1
#define RL_FUNC_ADD 0x8006
static inline void rlBegin(int m) { (void)m; }
static inline void rlEnd(void) {}
static inline void rlSetTexture(unsigned int id) { (void)id; }
static inline void rlColor4ub(unsigned char r, unsigned char g, unsigned char b, unsigned char a) { (void)r; (void)g; (void)b; (void)a; }
static inline void rlTexCoord2f(float x, float y) { (void)x; (void)y; }
static inline void rlVertex3f(float x, float y, float z) { (void)x; (void)y; (void)z; }
static inline void rlNormal3f(float x, float y, float z) { (void)x; (void)y; (void)z; }
static inline void rlDisableBackfaceCulling(void) {}
static inline void rlEnableBackfaceCulling(void) {}
static inline void rlDisableDepthMask(void) {}
static inline void rlEnableDepthMask(void) {}
static inline void rlDisableDepthTest(void) {}
static inline void rlEnableDepthTest(void) {}
static inline void rlSetBlendMode(int m) { (void)m; }
static inline void rlSetBlendFactors(int a, int b, int c) { (void)a; (void)b; (void)c; }
static inline void rlDrawRenderBatchActive(void) {}
static inline void rlPushMatrix(void) {}
static inline void rlEnableColorBlend(void) {}
static inline void rlDisableColorBlend(void) {}
static inline void rlPopMatrix(void) {}
#endif
