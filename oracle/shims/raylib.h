/* Fake "raylib.h" — TEST INFRASTRUCTURE ONLY.
 * The reference's CPU/3-D solvers include raylib for their window; the image has no raylib.  This
 * stub declares just enough (as no-ops) for those translation units to compile when they are
 * #included with `-Dmain=ref_main`; no drawing function is ever reached.  Written from the symbol
 * list in SURVEY.md §8(c); contains no reference code. */
#ifndef TAU_FAKE_RAYLIB_H
#define TAU_FAKE_RAYLIB_H
#include <stdbool.h>
typedef struct Color { unsigned char r, g, b, a; } Color;
typedef struct Rectangle { float x, y, width, height; } Rectangle;
typedef struct Vector2 { float x, y; } Vector2;
typedef struct Vector3 { float x, y, z; } Vector3;
typedef struct Image { void *data; int width, height, mipmaps, format; } Image;
typedef struct Texture2D { unsigned int id; int width, height, mipmaps, format; } Texture2D;
typedef struct Camera3D { Vector3 position, target, up; float fovy; int projection; } Camera3D;
typedef Camera3D Camera;
#define PIXELFORMAT_UNCOMPRESSED_R8G8B8A8 7
#define CAMERA_PERSPECTIVE 0
#define KEY_SPACE 32
#define KEY_R 82
#define KEY_M 77
#define KEY_ONE 49
#define KEY_TWO 50
#define KEY_THREE 51
#define KEY_FOUR 52
#define KEY_FIVE 53
#define KEY_SIX 54
#define KEY_SEVEN 55
#define KEY_EIGHT 56
#define KEY_UP 265
#define KEY_DOWN 264
#define KEY_LEFT 263
#define KEY_RIGHT 262
#define MOUSE_BUTTON_LEFT 0
#define MOUSE_BUTTON_RIGHT 1
#define MOUSE_BUTTON_MIDDLE 2
#define KEY_L 76
#define KEY_MINUS 45
#define KEY_EQUAL 61
#define KEY_KP_SUBTRACT 333
#define KEY_KP_ADD 334
#define KEY_LEFT_BRACKET 91
#define KEY_RIGHT_BRACKET 93
#define KEY_P 80
#define KEY_V 86
#define KEY_TAB 258
#define KEY_ESCAPE 256
#define BLACK ((Color){0, 0, 0, 255})
#define WHITE ((Color){255, 255, 255, 255})
#define GREEN ((Color){0, 228, 48, 255})
#define RAYWHITE ((Color){245, 245, 245, 255})
#define BLANK ((Color){0, 0, 0, 0})
#ifdef __cplusplus
#undef BLACK
#undef WHITE
#undef GREEN
#undef RAYWHITE
#undef BLANK
#define BLACK (Color{0, 0, 0, 255})
#define WHITE (Color{255, 255, 255, 255})
#define GREEN (Color{0, 228, 48, 255})
#define RAYWHITE (Color{245, 245, 245, 255})
#define BLANK (Color{0, 0, 0, 0})
extern "C" {
#endif
static inline void InitWindow(int w, int h, const char *t) { (void)w; (void)h; (void)t; }
static inline void SetTargetFPS(int f) { (void)f; }
static inline Texture2D LoadTextureFromImage(Image i) { Texture2D t = {0, i.width, i.height, 1, i.format}; return t; }
static inline void UnloadTexture(Texture2D t) { (void)t; }
static inline bool WindowShouldClose(void) { return true; }
static inline bool IsKeyPressed(int k) { (void)k; return false; }
static inline bool IsKeyDown(int k) { (void)k; return false; }
static inline bool IsMouseButtonDown(int b) { (void)b; return false; }
static inline Vector2 GetMouseDelta(void) { Vector2 v = {0, 0}; return v; }
static inline float GetMouseWheelMove(void) { return 0.0f; }
static inline float GetFrameTime(void) { return 0.0f; }
static inline int GetFPS(void) { return 0; }
static inline void UpdateTexture(Texture2D t, const void *p) { (void)t; (void)p; }
static inline void BeginDrawing(void) {}
static inline void EndDrawing(void) {}
static inline void BeginMode3D(Camera3D c) { (void)c; }
static inline void EndMode3D(void) {}
static inline void ClearBackground(Color c) { (void)c; }
static inline void DrawTexturePro(Texture2D t, Rectangle s, Rectangle d, Vector2 o, float r, Color c) { (void)t; (void)s; (void)d; (void)o; (void)r; (void)c; }
static inline void DrawText(const char *t, int x, int y, int s, Color c) { (void)t; (void)x; (void)y; (void)s; (void)c; }
static inline void DrawFPS(int x, int y) { (void)x; (void)y; }
static inline const char *TextFormat(const char *f, ...) { return f; }
static inline void CloseWindow(void) {}
static inline void DrawCubeWires(Vector3 p, float w, float h, float l, Color c) { (void)p; (void)w; (void)h; (void)l; (void)c; }
static inline Color Fade(Color c, float a) { (void)a; return c; }
static inline void DrawRectangle(int x, int y, int w, int h, Color c) { (void)x; (void)y; (void)w; (void)h; (void)c; }
#ifdef __cplusplus
}
#endif
#endif
