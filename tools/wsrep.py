"""whitespace-insensitive exact replacement helper for source edits: rep(text, old, new) replaces the single
occurrence of `old` (any whitespace run matches any whitespace run) by `new`"""
import re


def rep(s, old, new, count=1):
    toks = old.split()
    pat = r'\s+'.join(re.escape(t) for t in toks)
    m = list(re.finditer(pat, s))
    assert len(m) == count, (len(m), old[:80])
    for mm in reversed(m):
        s = s[:mm.start()] + new.strip('\n').lstrip(' \t') + s[mm.end():]
    return s
