/* TEST INFRASTRUCTURE ONLY (oracle/).  Minimal stand-in for the seven htslib
 * bgzf_* symbols the reference's libStatGen layer links against
 * (declared in VerifyBamID/statgen/bgzf.h of the reference tree; htslib itself
 * is not available offline).  Streams go through plain zlib gz* calls, so a
 * ".bam" written by FASTQuick_ref is a single gzip member holding the BAM
 * byte stream: compare *records*, never file bytes.                       */
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "bgzf.h"

static BGZF *wrap(gzFile g, int writing)
{
    BGZF *h;
    if (!g) return 0;
    h = (BGZF *)calloc(1, sizeof(BGZF));
    h->open_mode = writing ? 'w' : 'r';
    h->fp = (void *)g;
    return h;
}
static const char *zmode(const char *mode, int *writing)
{
    *writing = (strchr(mode, 'w') != 0) && (strchr(mode, 'r') == 0);
    return *writing ? "wb1" : "rb";
}
BGZF *bgzf_open(const char *path, const char *mode)
{
    int w; const char *m = zmode(mode, &w);
    return wrap(gzopen(path, m), w);
}
BGZF *bgzf_dopen(int fd, const char *mode)
{
    int w; const char *m = zmode(mode, &w);
    return wrap(gzdopen(fd, m), w);
}
int bgzf_close(BGZF *fp)
{
    int r;
    if (!fp) return -1;
    r = gzclose((gzFile)fp->fp);
    free(fp);
    return r == Z_OK ? 0 : -1;
}
ssize_t bgzf_read(BGZF *fp, void *data, ssize_t length)
{
    int n = gzread((gzFile)fp->fp, data, (unsigned)length);
    if (n > 0) fp->block_address += n;
    return n;
}
ssize_t bgzf_write(BGZF *fp, const void *data, ssize_t length)
{
    int n = gzwrite((gzFile)fp->fp, data, (unsigned)length);
    if (n > 0) fp->block_address += n;
    return n;
}
int64_t bgzf_seek(BGZF *fp, int64_t pos, int whence)
{
    (void)whence;
    if (gzseek((gzFile)fp->fp, (z_off_t)(pos >> 16), SEEK_SET) < 0) return -1;
    fp->block_address = pos >> 16; fp->block_offset = 0;
    return 0;
}
int bgzf_check_EOF(BGZF *fp) { (void)fp; return 1; }
