#!/usr/bin/env python3
"""One-off source transform used to introduce the `real` typedef (Float32 build): in the CODE part of every line (not in
// comments, not in string literals) `double` becomes `real` and every floating literal x becomes RL(x) = ((real)(x)).
Lines carrying the marker /*f64*/ and regions between /*f64-begin*/ and /*f64-end*/ are left alone.  The Float64 build must come out bit-identical (SASS compared)."""
import re
import sys

LIT = re.compile(r'(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])')


def conv_code(code):
    # leave string literals alone
    parts = re.split(r'("(?:[^"\\]|\\.)*")', code)
    for n in range(0, len(parts), 2):
        p = parts[n]
        p = re.sub(r'\bdouble\b', 'real', p)
        p = LIT.sub(lambda m: 'RL(%s)' % m.group(1), p)
        parts[n] = p
    return ''.join(parts)


def conv(text):
    out = []
    in_block = False
    keep = False      # between /*f64-begin*/ and /*f64-end*/ nothing is touched
    for line in text.split('\n'):
        if '/*f64-begin*/' in line:
            keep = True
        if '/*f64-end*/' in line:
            keep = False
            out.append(line)
            continue
        if keep or '/*f64*/' in line or line.lstrip().startswith('#include'):
            out.append(line)
            continue
        if in_block:
            if '*/' in line:
                in_block = False
            out.append(line)
            continue
        if line.lstrip().startswith('/*') and '*/' not in line:
            in_block = True
            out.append(line)
            continue
        i = line.find('//')
        code, comment = (line, '') if i < 0 else (line[:i], line[i:])
        out.append(conv_code(code) + comment)
    return '\n'.join(out)


if __name__ == '__main__':
    for p in sys.argv[1:]:
        s = open(p).read()
        open(p, 'w').write(conv(s))
        print('converted', p)
