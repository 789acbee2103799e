/*
 * config.c -- ns_config: layered INI / JSON configuration (host C).
 *
 * Behavioural mirror of /root/reference/src/common/config_parser.c behind the same
 * header (include/ns_config.h):
 *   - flat store of "section.key" -> string, later ns_config_set overrides (:78-105)
 *   - getters fall back to the default on a missing or unparsable value (:127-192);
 *     booleans accept true/false/yes/no/on/off (any case) and 1/0
 *   - INI: [section] headers, key = value, '#' / ';' full-line comments (:205-263)
 *   - JSON: nested objects flatten to dotted keys; strings, numbers (re-printed with
 *     %.15g), true/false, null -> ""; no arrays (:340-486)
 *   - a path ending in ".json" (any case) is JSON, anything else INI (:488-541)
 *   - schema validation with type / min / max / allowed-set checks (:556-634)
 * Written from the behaviour, not the text, of the reference.
 */
#include "ns_config.h"

#include <ctype.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { KEY_MAX = 256, VALUE_MAX = 512, LINE_MAX_LEN = 1024 };

/* ------------------------------------------------------------------ helpers */

static char *dup_string(const char *s) {
    const size_t n = strlen(s) + 1;
    char *d = (char *)malloc(n);
    if (d) memcpy(d, s, n);
    return d;
}

static void report(char *err, size_t cap, const char *what, const char *detail) {
    if (!err || cap == 0) return;
    if (detail) snprintf(err, cap, "%s: %s", what, detail);
    else snprintf(err, cap, "%s", what);
}

static int equal_nocase(const char *a, const char *b) {
    for (; *a && *b; ++a, ++b)
        if (tolower((unsigned char)*a) != tolower((unsigned char)*b)) return 0;
    return *a == '\0' && *b == '\0';
}

/* strips leading/trailing whitespace in place, returns s */
static char *strip(char *s) {
    char *b = s;
    while (*b && isspace((unsigned char)*b)) ++b;
    char *e = b + strlen(b);
    while (e > b && isspace((unsigned char)e[-1])) --e;
    *e = '\0';
    if (b != s) memmove(s, b, (size_t)(e - b) + 1);
    return s;
}

static NSConfigEntry *lookup(const NSConfig *cfg, const char *key) {
    if (!cfg || !key) return NULL;
    for (size_t k = 0; k < cfg->count; ++k)
        if (strcmp(cfg->entries[k].key, key) == 0) return &cfg->entries[k];
    return NULL;
}

static int parse_long(const char *s, long *out) {
    if (!s || !*s) return 0;
    char *end = NULL;
    const long v = strtol(s, &end, 10);
    if (end && *end != '\0') return 0;
    *out = v;
    return 1;
}

static int parse_double(const char *s, double *out) {
    if (!s || !*s) return 0;
    char *end = NULL;
    const double v = strtod(s, &end);
    if (end && *end != '\0') return 0;
    *out = v;
    return 1;
}

/* 1 true, 0 false, -1 not a boolean */
static int parse_bool(const char *s) {
    if (!s || !*s) return -1;
    if (equal_nocase(s, "true") || equal_nocase(s, "yes") || equal_nocase(s, "on") || strcmp(s, "1") == 0) return 1;
    if (equal_nocase(s, "false") || equal_nocase(s, "no") || equal_nocase(s, "off") || strcmp(s, "0") == 0) return 0;
    return -1;
}

/* -------------------------------------------------------------------- store */

void ns_config_init(NSConfig *cfg) {
    if (!cfg) return;
    cfg->entries = NULL;
    cfg->count = 0;
    cfg->capacity = 0;
}

void ns_config_free(NSConfig *cfg) {
    if (!cfg) return;
    for (size_t k = 0; k < cfg->count; ++k) {
        free(cfg->entries[k].key);
        free(cfg->entries[k].value);
    }
    free(cfg->entries);
    ns_config_init(cfg);
}

int ns_config_set(NSConfig *cfg, const char *key, const char *value) {
    if (!cfg || !key || !value) return -1;
    NSConfigEntry *hit = lookup(cfg, key);
    if (hit) {
        char *v = dup_string(value);
        if (!v) return -1;
        free(hit->value);
        hit->value = v;
        return 0;
    }
    if (cfg->count == cfg->capacity) {
        const size_t cap = cfg->capacity ? cfg->capacity * 2 : 16;
        NSConfigEntry *grown = (NSConfigEntry *)realloc(cfg->entries, cap * sizeof(*grown));
        if (!grown) return -1;
        cfg->entries = grown;
        cfg->capacity = cap;
    }
    char *k = dup_string(key), *v = dup_string(value);
    if (!k || !v) {
        free(k);
        free(v);
        return -1;
    }
    cfg->entries[cfg->count].key = k;
    cfg->entries[cfg->count].value = v;
    cfg->count++;
    return 0;
}

int ns_config_has_key(const NSConfig *cfg, const char *key) { return lookup(cfg, key) != NULL; }

const char *ns_config_get_string(const NSConfig *cfg, const char *key, const char *fallback) {
    const NSConfigEntry *e = lookup(cfg, key);
    return e ? e->value : fallback;
}

int ns_config_get_int(const NSConfig *cfg, const char *key, int fallback) {
    const NSConfigEntry *e = lookup(cfg, key);
    long v;
    return (e && parse_long(e->value, &v)) ? (int)v : fallback;
}

double ns_config_get_double(const NSConfig *cfg, const char *key, double fallback) {
    const NSConfigEntry *e = lookup(cfg, key);
    double v;
    return (e && parse_double(e->value, &v)) ? v : fallback;
}

int ns_config_get_bool(const NSConfig *cfg, const char *key, int fallback) {
    const NSConfigEntry *e = lookup(cfg, key);
    if (!e) return fallback;
    const int b = parse_bool(e->value);
    return b < 0 ? fallback : b;
}

/* ---------------------------------------------------------------------- INI */

static int load_ini(NSConfig *cfg, FILE *fp, char *err, size_t cap) {
    char section[KEY_MAX] = "";
    char line[LINE_MAX_LEN];
    while (fgets(line, sizeof line, fp)) {
        char *s = strip(line);
        if (*s == '\0' || *s == '#' || *s == ';') continue;
        if (*s == '[') {
            char *close = strchr(s, ']');
            if (!close) {
                report(err, cap, "Malformed section header in INI file", s);
                return -1;
            }
            *close = '\0';
            snprintf(section, sizeof section, "%s", s + 1);
            strip(section);
            continue;
        }
        char *eq = strchr(s, '=');
        if (!eq) {
            report(err, cap, "Malformed key-value line in INI file", s);
            return -1;
        }
        *eq = '\0';
        char name[KEY_MAX], value[VALUE_MAX], full[KEY_MAX];
        snprintf(name, sizeof name, "%s", s);
        snprintf(value, sizeof value, "%s", eq + 1);
        strip(name);
        strip(value);
        int n;
        if (section[0]) n = snprintf(full, sizeof full, "%s.%s", section, name);
        else n = snprintf(full, sizeof full, "%s", name);
        if (section[0] && (n < 0 || (size_t)n >= sizeof full)) {
            report(err, cap, "Combined key too long", name);
            return -1;
        }
        if (ns_config_set(cfg, full, value) != 0) {
            report(err, cap, "Failed to store configuration key", full);
            return -1;
        }
    }
    return 0;
}

/* --------------------------------------------------------------------- JSON */

typedef struct {
    const char *p;
    NSConfig *cfg;
    char *err;
    size_t cap;
} JsonCursor;

static void skip_space(JsonCursor *c) {
    while (*c->p && isspace((unsigned char)*c->p)) ++c->p;
}

/* returns a malloc'd string, cursor past the closing quote; NULL on error */
static char *json_string(JsonCursor *c) {
    skip_space(c);
    if (*c->p != '"') {
        report(c->err, c->cap, "Expected string in JSON configuration", NULL);
        return NULL;
    }
    ++c->p;
    size_t cap = 64, len = 0;
    char *buf = (char *)malloc(cap);
    if (!buf) {
        report(c->err, c->cap, "Out of memory parsing JSON string", NULL);
        return NULL;
    }
    while (*c->p && *c->p != '"') {
        char ch = *c->p;
        if (ch == '\\') {
            ++c->p;
            if (!*c->p) {
                free(buf);
                report(c->err, c->cap, "Incomplete escape sequence in JSON string", NULL);
                return NULL;
            }
            switch (*c->p) {
                case 'b': ch = '\b'; break;
                case 'f': ch = '\f'; break;
                case 'n': ch = '\n'; break;
                case 'r': ch = '\r'; break;
                case 't': ch = '\t'; break;
                default: ch = *c->p; break; /* \" \\ \/ and anything else: the character itself */
            }
        }
        if (len + 2 > cap) {
            cap *= 2;
            char *grown = (char *)realloc(buf, cap);
            if (!grown) {
                free(buf);
                report(c->err, c->cap, "Out of memory expanding JSON string", NULL);
                return NULL;
            }
            buf = grown;
        }
        buf[len++] = ch;
        ++c->p;
    }
    if (*c->p != '"') {
        free(buf);
        report(c->err, c->cap, "Unterminated string in JSON configuration", NULL);
        return NULL;
    }
    ++c->p;
    buf[len] = '\0';
    return buf;
}

static int json_object(JsonCursor *c, const char *prefix);

static int json_store(JsonCursor *c, const char *key, const char *value, const char *what) {
    if (ns_config_set(c->cfg, key, value) != 0) {
        report(c->err, c->cap, what, key);
        return -1;
    }
    return 0;
}

static int json_value(JsonCursor *c, const char *key) {
    skip_space(c);
    if (!*c->p) {
        report(c->err, c->cap, "Unexpected end of JSON value", NULL);
        return -1;
    }
    if (*c->p == '{') return json_object(c, key);
    if (*c->p == '"') {
        char *s = json_string(c);
        if (!s) return -1;
        const int rc = json_store(c, key, s, "Failed to store JSON string value");
        free(s);
        return rc;
    }
    if (strncmp(c->p, "true", 4) == 0) { c->p += 4; return json_store(c, key, "true", "Failed to store JSON boolean"); }
    if (strncmp(c->p, "false", 5) == 0) { c->p += 5; return json_store(c, key, "false", "Failed to store JSON boolean"); }
    if (strncmp(c->p, "null", 4) == 0) { c->p += 4; return json_store(c, key, "", "Failed to store JSON null"); }
    char *end = NULL;
    const double number = strtod(c->p, &end);
    if (end == c->p) {
        report(c->err, c->cap, "Invalid JSON value", key);
        return -1;
    }
    char text[64];
    snprintf(text, sizeof text, "%.15g", number);
    c->p = end;
    return json_store(c, key, text, "Failed to store JSON numeric value");
}

static int json_object(JsonCursor *c, const char *prefix) {
    skip_space(c);
    if (*c->p != '{') {
        report(c->err, c->cap, "Expected '{' in JSON configuration", NULL);
        return -1;
    }
    ++c->p;
    skip_space(c);
    if (*c->p == '}') { ++c->p; return 0; }
    while (*c->p) {
        char *name = json_string(c);
        if (!name) return -1;
        skip_space(c);
        if (*c->p != ':') {
            free(name);
            report(c->err, c->cap, "Expected ':' in JSON key/value pair", NULL);
            return -1;
        }
        ++c->p;
        char full[KEY_MAX];
        int n;
        if (prefix && prefix[0]) n = snprintf(full, sizeof full, "%s.%s", prefix, name);
        else n = snprintf(full, sizeof full, "%s", name);
        free(name);
        if (prefix && prefix[0] && (n < 0 || (size_t)n >= sizeof full)) {
            report(c->err, c->cap, "Combined JSON key too long", NULL);
            return -1;
        }
        if (json_value(c, full) != 0) return -1;
        skip_space(c);
        if (*c->p == ',') { ++c->p; continue; }
        if (*c->p == '}') { ++c->p; return 0; }
        report(c->err, c->cap, "Expected ',' or '}' in JSON object", NULL);
        return -1;
    }
    report(c->err, c->cap, "Unexpected end of JSON content", NULL);
    return -1;
}

static int load_json_text(NSConfig *cfg, const char *text, char *err, size_t cap) {
    JsonCursor c = {text, cfg, err, cap};
    if (json_object(&c, "") != 0) return -1;
    skip_space(&c);
    if (*c.p != '\0') {
        report(err, cap, "Unexpected trailing data in JSON file", NULL);
        return -1;
    }
    return 0;
}

/* -------------------------------------------------------------------- files */

static int load_one(NSConfig *cfg, const char *path, char *err, size_t cap) {
    const char *dot = strrchr(path, '.');
    if (dot && equal_nocase(dot, ".json")) {
        FILE *fp = fopen(path, "rb");
        if (!fp) {
            report(err, cap, "Failed to open JSON configuration file", strerror(errno));
            return -1;
        }
        long size = -1;
        if (fseek(fp, 0, SEEK_END) == 0) size = ftell(fp);
        if (size < 0 || fseek(fp, 0, SEEK_SET) != 0) {
            fclose(fp);
            report(err, cap, "Failed to determine JSON file size", path);
            return -1;
        }
        char *text = (char *)malloc((size_t)size + 1);
        if (!text) {
            fclose(fp);
            report(err, cap, "Out of memory reading JSON configuration file", path);
            return -1;
        }
        const size_t got = fread(text, 1, (size_t)size, fp);
        fclose(fp);
        if (got != (size_t)size) {
            free(text);
            report(err, cap, "Failed to read JSON configuration file", path);
            return -1;
        }
        text[size] = '\0';
        const int rc = load_json_text(cfg, text, err, cap);
        free(text);
        return rc;
    }
    FILE *fp = fopen(path, "r");
    if (!fp) {
        report(err, cap, "Failed to open INI configuration file", strerror(errno));
        return -1;
    }
    const int rc = load_ini(cfg, fp, err, cap);
    fclose(fp);
    return rc;
}

int ns_config_load_files(NSConfig *cfg, const char **paths, size_t count, char *err, size_t cap) {
    if (!cfg || (!paths && count > 0)) {
        report(err, cap, "Invalid arguments to ns_config_load_files", NULL);
        return -1;
    }
    for (size_t k = 0; k < count; ++k)
        if (load_one(cfg, paths[k], err, cap) != 0) return -1;
    return 0;
}

/* --------------------------------------------------------------- validation */

int ns_config_validate(const NSConfig *cfg, const NSConfigSchemaEntry *schema, size_t n, char *err, size_t cap) {
    if (!cfg || !schema) {
        report(err, cap, "Invalid arguments to ns_config_validate", NULL);
        return -1;
    }
    for (size_t k = 0; k < n; ++k) {
        const NSConfigSchemaEntry *s = &schema[k];
        const NSConfigEntry *e = lookup(cfg, s->key);
        if (!e) {
            if (s->required) {
                report(err, cap, "Missing required configuration key", s->key);
                return -1;
            }
            continue;
        }
        const char *v = e->value;
        switch (s->type) {
            case NS_CONFIG_VALUE_STRING: {
                if (!s->allowed_values || s->allowed_count == 0) break;
                int ok = 0;
                for (size_t a = 0; a < s->allowed_count && !ok; ++a) ok = equal_nocase(v, s->allowed_values[a]);
                if (!ok) {
                    report(err, cap, "Configuration string value not permitted", s->key);
                    return -1;
                }
                break;
            }
            case NS_CONFIG_VALUE_INT: {
                long iv;
                if (!parse_long(v, &iv)) {
                    report(err, cap, "Configuration key expects integer value", s->key);
                    return -1;
                }
                if (s->has_min && iv < s->min_value) {
                    report(err, cap, "Configuration integer below minimum", s->key);
                    return -1;
                }
                if (s->has_max && iv > s->max_value) {
                    report(err, cap, "Configuration integer above maximum", s->key);
                    return -1;
                }
                break;
            }
            case NS_CONFIG_VALUE_DOUBLE: {
                double dv;
                if (!parse_double(v, &dv)) {
                    report(err, cap, "Configuration key expects numeric value", s->key);
                    return -1;
                }
                if (s->has_min && dv < s->min_value) {
                    report(err, cap, "Configuration value below minimum", s->key);
                    return -1;
                }
                if (s->has_max && dv > s->max_value) {
                    report(err, cap, "Configuration value above maximum", s->key);
                    return -1;
                }
                break;
            }
            case NS_CONFIG_VALUE_BOOL:
                if (parse_bool(v) < 0) {
                    report(err, cap, "Configuration key expects boolean value", s->key);
                    return -1;
                }
                break;
        }
    }
    return 0;
}
