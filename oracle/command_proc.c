/*
 * command_proc.c -- ORACLE (test infrastructure).  What amps.command_processor publishes for one text command:
 * restates lib/command_processor_impl.cc:52-117 (commands_message, handle_page, debug_msg).
 *
 * Matching is by prefix, in the reference's order: "fvc off", "fvc on", "fvc alert" are case-SENSITIVE
 * (boost::starts_with), "page " is case-INSENSITIVE (boost::istarts_with); the page argument is trimmed of
 * white space on both ends (boost::trim).  A MIN shorter than 10 digits makes the reference read past the end
 * of the string (lib/amps_packet.h:338-346); like orc_parse_min this restatement rejects it ("invalid MIN entered").
 */
#include "amps_oracle.h"
#include <ctype.h>
#include <string.h>

static int starts_with(const char *s, const char *p) { return strncmp(s, p, strlen(p)) == 0; }
static int istarts_with(const char *s, const char *p) {
    for (; *p; s++, p++) if (tolower((unsigned char)*s) != tolower((unsigned char)*p)) return 0;
    return 1;
}

void orc_command_actions(const char *cmd, orc_cmd_actions *a) {
    memset(a, 0, sizeof *a);
    a->fvc_mute = -1; a->audio_mute = -1;
    if (starts_with(cmd, "fvc off")) {                               /* :93-96 */
        a->fvc_mute = 1; a->audio_mute = 0;
        strcpy(a->debug[a->n_debug++], "turning FVC data OFF; audio ON\n");
    } else if (starts_with(cmd, "fvc on")) {                         /* :97-100 */
        a->fvc_mute = 0; a->audio_mute = 1;
        strcpy(a->debug[a->n_debug++], "turning FVC data ON; audio OFF\n");
    } else if (starts_with(cmd, "fvc alert")) {                      /* :101-105: tuple (1, word) -- no timer element */
        a->has_fvc = 1;
        orc_fvc_word1_general(a->fvc_word, 1, 0, 0, 1);
    } else if (istarts_with(cmd, "page ")) {                         /* :106-109 + handle_page :58-82 */
        const char *b = cmd + 5, *e = cmd + strlen(cmd);
        while (b < e && isspace((unsigned char)*b)) b++;
        while (e > b && isspace((unsigned char)e[-1])) e--;
        char num[64];
        size_t n = (size_t)(e - b);
        if (n >= sizeof num) n = sizeof num - 1;
        memcpy(num, b, n); num[n] = 0;
        if (n < 1) { strcpy(a->debug[a->n_debug++], "missing MIN in page command\n"); return; }
        strcpy(a->debug[a->n_debug++], "paging!\n");
        uint64_t m1, m2;
        if (!orc_parse_min(num, &m1, &m2)) { strcpy(a->debug[a->n_debug++], "invalid MIN entered"); return; }
        a->n_focc = 2; a->focc_stream = 3;                           /* page = Word 1 + Word 2 with order 0 */
        orc_focc_word1(a->focc_words[0], 1, 0, m1);
        orc_focc_word2_general(a->focc_words[1], m2, 0, 0, 0);
    } else {
        strcpy(a->debug[a->n_debug++], "invalid command\n");         /* :110-112 */
    }
}
