/*
 * coast_scope.c -- the front end of the BOARD=b200 "pass": which functions are inside the sphere of replication?
 *
 * The reference decides that inside `opt`: processAnnotations() reads llvm.global.annotations
 * (projects/dataflowProtection/interface.cpp:364-532: "xMR" -> fnsToClone, "no_xMR" -> fnsToSkip,
 * "set_no_xMR_default" -> xMR_default = false) and everything not excluded is cloned when xMR_default is true
 * (dataflowProtection.h:61-62).  There is no LLVM here, so the same decision is taken from the C sources:
 *
 *   scan : stdin = `gcc -E -DCOAST_SCOPE_SCAN` output, in which every COAST.h directive survives as the token
 *          sequence  __coast_anno__("<annotation string>")  with the reference's own strings
 *          (dataflowProtection.h:69-79).  stdout = one fact per line:
 *              default xMR|no_xMR        fn <name> <annotation>        var <name> <annotation>
 *              def <name>                call <caller> <callee>        local <function> <annotation>
 *   plan : reads the facts of every translation unit of the program (the reference links the bitcode before `opt`,
 *          Makefile.compile.x86, so scope is per PROGRAM) + the runtime's entry table, and writes the sed script the
 *          pass applies to the assembly.  Per function F that is in the SoR:
 *              F has a runtime kernel                      -> calls to F are redirected to the runtime entry
 *              F is computed inside a kernel (subsumed)    -> nothing to do
 *              F reaches a kernel entry through its calls  -> host wrapper: its protected work is that call
 *              F is EXPLICITLY __xMR and none of the above -> the build FAILS, naming F   (COAST_HOST_OK lists the
 *                                                             reference's own harness functions that are allowed to
 *                                                             run unprotected, with a warning)
 *          A function the program marks __NO_xMR -- or leaves out of the SoR under __DEFAULT_NO_xMR -- is never
 *          redirected: it runs once, unprotected, on the CPU, as in the reference.
 *
 * Plain C, no dependencies; built by Makefile.common next to the objects.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ tokens */
enum { T_EOF, T_ID, T_STR, T_NUM, T_PUNCT };
typedef struct { int kind; char text[256]; } token;

static int next_token(FILE* f, token* t) {
    int c;
    for (;;) {
        c = fgetc(f);
        if (c == EOF) { t->kind = T_EOF; t->text[0] = 0; return 0; }
        if (c == '#') { while ((c = fgetc(f)) != EOF && c != '\n') {} continue; }      /* line markers / pragmas */
        if (!isspace(c)) break;
    }
    size_t n = 0;
    if (isalpha(c) || c == '_') {
        while (c != EOF && (isalnum(c) || c == '_')) { if (n < sizeof t->text - 1) t->text[n++] = (char)c; c = fgetc(f); }
        if (c != EOF) ungetc(c, f);
        t->kind = T_ID;
    } else if (isdigit(c)) {
        while (c != EOF && (isalnum(c) || c == '.' || c == '_')) { if (n < sizeof t->text - 1) t->text[n++] = (char)c; c = fgetc(f); }
        if (c != EOF) ungetc(c, f);
        t->kind = T_NUM;
    } else if (c == '"' || c == '\'') {
        int q = c;
        while ((c = fgetc(f)) != EOF && c != q) {
            if (c == '\\') { int e = fgetc(f); if (n < sizeof t->text - 2) { t->text[n++] = '\\'; t->text[n++] = (char)e; } continue; }
            if (n < sizeof t->text - 1) t->text[n++] = (char)c;
        }
        t->kind = q == '"' ? T_STR : T_NUM;
    } else {
        t->text[n++] = (char)c;
        t->kind = T_PUNCT;
    }
    t->text[n] = 0;
    return 1;
}

static int is_keyword(const char* s) {
    static const char* const kw[] = { "if", "else", "while", "for", "do", "switch", "case", "return", "sizeof", "goto", "break",
        "continue", "default", "typedef", "struct", "union", "enum", "static", "extern", "inline", "const", "volatile", "register",
        "unsigned", "signed", "int", "char", "short", "long", "float", "double", "void", "_Bool", "__attribute__", "__attribute",
        "__coast_anno__", "__asm__", "asm", "__extension__", "__inline", "__inline__", "__restrict", "restrict", "__typeof__",
        "typeof", "_Alignas", "_Alignof", "__alignof__", "_Static_assert", "__builtin_va_arg", "__builtin_offsetof", NULL };
    for (int i = 0; kw[i]; ++i) if (!strcmp(kw[i], s)) return 1;
    return 0;
}

/* ------------------------------------------------------------------ scan */
#define MAX_HEAD 4096
static token head[MAX_HEAD];
static int n_head;

/* name of the entity an external declaration declares: *is_fn = identifier directly followed by '(' at depth 0 */
static const char* decl_name(int* is_fn) {
    int depth = 0;
    const char* last_id = NULL;
    *is_fn = 0;
    for (int i = 0; i < n_head; ++i) {
        const token* t = &head[i];
        if (t->kind == T_ID && (!strcmp(t->text, "__attribute__") || !strcmp(t->text, "__attribute") || !strcmp(t->text, "__coast_anno__") ||
                                !strcmp(t->text, "__asm__") || !strcmp(t->text, "asm"))) {
            int d = 0;                                               /* skip the balanced ( ... ) that follows */
            for (++i; i < n_head; ++i) {
                if (head[i].kind == T_PUNCT && head[i].text[0] == '(') ++d;
                else if (head[i].kind == T_PUNCT && head[i].text[0] == ')') { if (--d == 0) break; }
                else if (!d) { --i; break; }
            }
            continue;
        }
        if (t->kind == T_PUNCT) {
            char c = t->text[0];
            if (c == '(' || c == '[') ++depth;
            else if (c == ')' || c == ']') --depth;
            else if (depth == 0 && (c == '=' || c == ',')) break;
            continue;
        }
        if (t->kind == T_ID && depth == 0 && !is_keyword(t->text)) {
            if (i + 1 < n_head && head[i + 1].kind == T_PUNCT && head[i + 1].text[0] == '(') { *is_fn = 1; return t->text; }
            last_id = t->text;
        }
    }
    return last_id;
}

static void emit_head_annotations(void) {
    int is_fn = 0;
    const char* name = decl_name(&is_fn);
    for (int i = 0; i + 3 < n_head; ++i) {
        if (head[i].kind == T_ID && !strcmp(head[i].text, "__coast_anno__") && head[i + 2].kind == T_STR) {
            const char* a = head[i + 2].text;
            if (!strcmp(a, "set_xMR_default")) puts("default xMR");
            else if (!strcmp(a, "set_no_xMR_default")) puts("default no_xMR");
            else if (name) printf("%s %s %s\n", is_fn ? "fn" : "var", name, a);
        }
    }
}

static int scan(FILE* f) {
    token t;
    int depth = 0;                       /* brace depth */
    char cur_fn[256] = "";
    token prev, prev2; prev.kind = prev2.kind = T_EOF; prev.text[0] = prev2.text[0] = 0;
    int skip_braces = 0;                 /* inside an initializer / struct body at file scope */
    n_head = 0;
    while (next_token(f, &t)) {
        if (depth == 0 && !skip_braces) {
            if (t.kind == T_PUNCT && t.text[0] == ';') {
                if (n_head && !(head[0].kind == T_ID && !strcmp(head[0].text, "typedef"))) emit_head_annotations();
                n_head = 0;
            } else if (t.kind == T_PUNCT && t.text[0] == '{') {
                /* function body iff the declarator is a function and what precedes '{' closes it: ')' or a trailing directive */
                int is_fn = 0;
                const char* name = decl_name(&is_fn);
                int last = n_head - 1;
                int closes = last >= 0 && head[last].kind == T_PUNCT && head[last].text[0] == ')';
                if (is_fn && name && closes && !(head[0].kind == T_ID && !strcmp(head[0].text, "typedef"))) {
                    emit_head_annotations();
                    printf("def %s\n", name);
                    snprintf(cur_fn, sizeof cur_fn, "%s", name);
                    n_head = 0;
                    depth = 1;
                } else {
                    skip_braces = 1;     /* `= { ... }` or `struct S { ... }`: part of the same declaration */
                }
            } else if (n_head < MAX_HEAD) {
                head[n_head++] = t;
            }
        } else if (skip_braces) {
            if (t.kind == T_PUNCT && t.text[0] == '{') ++skip_braces;
            else if (t.kind == T_PUNCT && t.text[0] == '}') --skip_braces;
        } else {
            if (t.kind == T_PUNCT && t.text[0] == '{') ++depth;
            else if (t.kind == T_PUNCT && t.text[0] == '}') { if (--depth == 0) cur_fn[0] = 0; }
            else if (t.kind == T_PUNCT && t.text[0] == '(' && prev.kind == T_ID && !is_keyword(prev.text))
                printf("call %s %s\n", cur_fn, prev.text);
            else if (t.kind == T_STR && prev.kind == T_PUNCT && prev.text[0] == '(' && prev2.kind == T_ID && !strcmp(prev2.text, "__coast_anno__"))
                printf("local %s %s\n", cur_fn, t.text);    /* a directive on a local variable (interface.cpp:540-601) */
        }
        prev2 = prev;
        prev = t;
    }
    return 0;
}

/* ------------------------------------------------------------------ plan */
#define MAX_N 2048
typedef struct { char name[128]; int explicit_xmr, explicit_no, defined, ret_val, call_once, isr; } fn_rec;
static fn_rec fns[MAX_N]; static int n_fns;
static struct { char a[128], b[128]; } calls[8 * MAX_N]; static int n_calls;
static struct { char fn[128], sym[128]; char sub[32][64]; int n_sub; } entries[64]; static int n_entries;

static fn_rec* fn_get(const char* name) {
    for (int i = 0; i < n_fns; ++i) if (!strcmp(fns[i].name, name)) return &fns[i];
    if (n_fns >= MAX_N) { fprintf(stderr, "coast: too many functions\n"); exit(2); }
    memset(&fns[n_fns], 0, sizeof fns[0]);
    snprintf(fns[n_fns].name, sizeof fns[0].name, "%s", name);
    return &fns[n_fns++];
}
static int entry_of(const char* fn) { for (int i = 0; i < n_entries; ++i) if (!strcmp(entries[i].fn, fn)) return i; return -1; }
static int subsumed_by(const char* fn) {
    for (int i = 0; i < n_entries; ++i) for (int j = 0; j < entries[i].n_sub; ++j) if (!strcmp(entries[i].sub[j], fn)) return i;
    return -1;
}
static int in_list(const char* list, const char* name) {
    size_t n = strlen(name);
    for (const char* p = list; p && *p; ) {
        while (*p == ' ' || *p == ',') ++p;
        const char* e = p; while (*e && *e != ' ' && *e != ',') ++e;
        if ((size_t)(e - p) == n && !strncmp(p, name, n)) return 1;
        p = e;
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "scan")) return scan(stdin);
    if (argc < 5 || strcmp(argv[1], "plan")) {
        fprintf(stderr, "usage: coast_scope scan < file.i > file.scope\n"
                        "       coast_scope plan <entries.tab> <out.sed> <host-ok list> <verbose 0|1> file.scope...\n");
        return 2;
    }
    const char* tab = argv[2]; const char* sed_out = argv[3]; const char* host_ok = argv[4]; const int verbose = atoi(argv[5]);
    FILE* f = fopen(tab, "r");
    if (!f) { fprintf(stderr, "coast: cannot read the entry table %s\n", tab); return 2; }
    char line[1024];
    while (fgets(line, sizeof line, f)) {
        char* h = strchr(line, '#'); if (h) *h = 0;
        char* tok = strtok(line, " \t\r\n"); if (!tok) continue;
        snprintf(entries[n_entries].fn, sizeof entries[0].fn, "%s", tok);
        tok = strtok(NULL, " \t\r\n"); if (!tok) { fprintf(stderr, "coast: %s: entry '%s' has no runtime symbol\n", tab, entries[n_entries].fn); return 2; }
        snprintf(entries[n_entries].sym, sizeof entries[0].sym, "%s", tok);
        entries[n_entries].n_sub = 0;
        while ((tok = strtok(NULL, " \t\r\n")) && entries[n_entries].n_sub < 32)
            snprintf(entries[n_entries].sub[entries[n_entries].n_sub++], 64, "%s", tok);
        if (++n_entries >= 64) break;
    }
    fclose(f);
    int xmr_default = 1;                                     /* dataflowProtection.h:62 */
    for (int i = 6; i < argc; ++i) {
        f = fopen(argv[i], "r");
        if (!f) { fprintf(stderr, "coast: cannot read %s\n", argv[i]); return 2; }
        while (fgets(line, sizeof line, f)) {
            char a[128] = "", b[128] = "", c[128] = "";
            int n = sscanf(line, "%127s %127s %127s", a, b, c);
            if (n >= 2 && !strcmp(a, "default")) { if (!strcmp(b, "no_xMR")) xmr_default = 0; }
            else if (n >= 3 && !strcmp(a, "fn")) {
                fn_rec* r = fn_get(b);
                if (!strcmp(c, "xMR") || !strcmp(c, "protected_lib")) { r->explicit_xmr = 1; r->explicit_no = 0; }   /* interface.cpp:389-391, 467-471 */
                else if (!strcmp(c, "no_xMR")) { r->explicit_no = 1; r->explicit_xmr = 0; }                           /* :383-388 */
                else if (!strcmp(c, "repl_return_val")) r->ret_val = 1;                                               /* :464-466 */
                else if (!strcmp(c, "coast_call_once")) r->call_once = 1;                                             /* :396-399 */
                else if (!strcmp(c, "isr_function")) { r->isr = 1; r->explicit_no = 1; r->explicit_xmr = 0; }         /* :461-463: never cloned */
            }
            else if (n >= 2 && !strcmp(a, "def")) fn_get(b)->defined = 1;
            else if (n >= 3 && !strcmp(a, "call") && n_calls < 8 * MAX_N) {
                snprintf(calls[n_calls].a, 128, "%s", b); snprintf(calls[n_calls].b, 128, "%s", c); ++n_calls;
            }
        }
        fclose(f);
    }
    /* reach[i]: function i reaches a kernel entry that is in the SoR through its calls (fixed point) */
    static int reach[MAX_N], in_sor[MAX_N];
    for (int k = 0; k < n_calls; ++k) { fn_get(calls[k].a); fn_get(calls[k].b); }          /* every endpoint has a record */
    for (int i = 0; i < n_fns; ++i) in_sor[i] = fns[i].explicit_xmr || (xmr_default && !fns[i].explicit_no);
    for (int changed = 1; changed; ) {
        changed = 0;
        for (int k = 0; k < n_calls; ++k) {
            const int ia = (int)(fn_get(calls[k].a) - fns), ib = (int)(fn_get(calls[k].b) - fns);
            if (reach[ia]) continue;
            if ((entry_of(calls[k].b) >= 0 && in_sor[ib]) || reach[ib]) { reach[ia] = 1; changed = 1; }
        }
    }
    FILE* sed = fopen(sed_out, "w");
    if (!sed) { fprintf(stderr, "coast: cannot write %s\n", sed_out); return 2; }
    fprintf(sed, "# generated by coast_scope plan: calls into the sphere of replication -> runtime launches\n");
    printf("  COAST   scope: default %s (%s)\n", xmr_default ? "xMR" : "no_xMR", xmr_default ? "no __DEFAULT_NO_xMR in the program" : "__DEFAULT_NO_xMR");
    int errors = 0, offloaded = 0;
    for (int i = 0; i < n_fns; ++i) {
        fn_rec* r = &fns[i];
        if (!r->defined) continue;
        const int e = entry_of(r->name), s = subsumed_by(r->name);
        const char* why = r->explicit_xmr ? "__xMR" : r->explicit_no ? "__NO_xMR" : xmr_default ? "default scope" : "outside the default scope";
        if (r->ret_val && (e >= 0 || verbose))
            printf("  COAST   ret-val   %s: __xMR_RET_VAL -- a kernel votes EVERY element it returns, field by field with the field's own width "
                   "(synchronization.cpp:816-913), so a replicated return value is what the runtime entry already delivers\n", r->name);
        if (e >= 0) {
            if (in_sor[i]) {
                /* `call crc16` / `jmp crc16` (with or without @PLT) -> the runtime entry */
                fprintf(sed, "s/\\(call\\|jmp\\)\\([[:space:]]\\+\\)%s\\(@PLT\\)\\?[[:space:]]*$/\\1\\2%s@PLT/\n", r->name, entries[e].sym);
                printf("  COAST   offload   %s -> %s   (%s)\n", r->name, entries[e].sym, why);
                ++offloaded;
            } else {
                printf("  COAST   cpu-only  %s stays on the CPU, unprotected (%s)\n", r->name, why);
            }
            continue;
        }
        if (!in_sor[i]) { if (verbose) printf("  COAST   cpu-only  %s (%s)\n", r->name, why); continue; }
        if (s >= 0) { if (verbose) printf("  COAST   inside    %s is computed inside the %s kernel\n", r->name, entries[s].fn); continue; }
        if (reach[i]) {
            if (r->explicit_xmr || verbose)
                printf("  COAST   wrapper   %s runs on the host; its protected work is the kernel entry it calls (%s)\n", r->name, why);
            continue;
        }
        if (r->explicit_xmr) {
            if (in_list(host_ok, r->name)) {
                printf("  COAST   WARNING   %s is marked __xMR but has no kernel: it runs UNPROTECTED on the host (allowed by COAST_HOST_OK)\n", r->name);
            } else {
                fprintf(stderr, "coast: error: function '%s' is marked __xMR but libcoast_rt has no protected kernel for it and it calls none.\n"
                                "coast:        kernel entries:", r->name);
                for (int k = 0; k < n_entries; ++k) fprintf(stderr, " %s", entries[k].fn);
                fprintf(stderr, "\ncoast:        Remove the directive, mark it __NO_xMR, or accept an unprotected host function with COAST_HOST_OK=%s\n", r->name);
                ++errors;
            }
        } else if (verbose) {
            printf("  COAST   host      %s (default scope, no kernel): runs unprotected on the host\n", r->name);
        }
    }
    fclose(sed);
    if (!offloaded && !errors) printf("  COAST   WARNING   nothing in this program is offloaded: no kernel entry is inside the sphere of replication\n");
    return errors ? 1 : 0;
}
