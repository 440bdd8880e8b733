/* Fake <ncursesw/curses.h> — TEST INFRASTRUCTURE ONLY.
 * The reference's tau_gray_scott.cu / tau_sph.cu / tau_burgers.cu include ncurses for their terminal renderer; the
 * image has no ncurses headers.  This stub declares just enough (as no-ops) for those translation
 * units to compile when oracle/ref_drivers/ #include them with `-Dmain=ref_main`; the renderer is
 * never called.  Written from the symbol list in SURVEY.md §8(c); contains no reference code. */
#ifndef TAU_FAKE_CURSES_H
#define TAU_FAKE_CURSES_H
#include <wchar.h>
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
#define ERR (-1)
#define OK 0
typedef struct tau_fake_window { int unused; } WINDOW;
static WINDOW tau_fake_stdscr_obj;
static WINDOW *stdscr = &tau_fake_stdscr_obj;
static int LINES = 24, COLS = 80;
#define KEY_UP 0403
#define KEY_DOWN 0402
#define KEY_LEFT 0404
#define KEY_RIGHT 0405
#define COLOR_BLACK 0
#define COLOR_RED 1
#define COLOR_GREEN 2
#define COLOR_YELLOW 3
#define COLOR_BLUE 4
#define COLOR_MAGENTA 5
#define COLOR_CYAN 6
#define COLOR_WHITE 7
#define COLOR_PAIR(n) (n)
#define A_BOLD 0
static inline WINDOW *initscr(void) { return stdscr; }
static inline int endwin(void) { return OK; }
static inline int noecho(void) { return OK; }
static inline int cbreak(void) { return OK; }
static inline int curs_set(int v) { (void)v; return OK; }
static inline int nodelay(WINDOW *w, int b) { (void)w; (void)b; return OK; }
static inline int keypad(WINDOW *w, int b) { (void)w; (void)b; return OK; }
static inline int has_colors(void) { return 0; }
static inline int start_color(void) { return OK; }
static inline int use_default_colors(void) { return OK; }
static inline int init_pair(short a, short b, short c) { (void)a; (void)b; (void)c; return OK; }
static inline int attron(int a) { (void)a; return OK; }
static inline int attroff(int a) { (void)a; return OK; }
static inline int erase(void) { return OK; }
static inline int clear(void) { return OK; }
static inline int move(int y, int x) { (void)y; (void)x; return OK; }
static inline int addwstr(const wchar_t *s) { (void)s; return OK; }
static inline int addstr(const char *s) { (void)s; return OK; }
static inline int addch(int c) { (void)c; return OK; }
static inline int mvaddch(int y, int x, int c) { (void)y; (void)x; (void)c; return OK; }
static inline int mvaddwstr(int y, int x, const wchar_t *s) { (void)y; (void)x; (void)s; return OK; }
static inline int mvaddstr(int y, int x, const char *s) { (void)y; (void)x; (void)s; return OK; }
static inline int mvprintw(int y, int x, const char *fmt, ...) { (void)y; (void)x; (void)fmt; return OK; }
static inline int printw(const char *fmt, ...) { (void)fmt; return OK; }
static inline int addnwstr(const wchar_t *s, int n) { (void)s; (void)n; return OK; }
static inline int clrtoeol(void) { return OK; }
static inline int refresh(void) { return OK; }
static inline int getch(void) { return ERR; }
static inline int timeout_(int t) { (void)t; return OK; }
#define timeout(t) timeout_(t)
#define getmaxyx(win, y, x) do { (void)(win); (y) = LINES; (x) = COLS; } while (0)
#endif
