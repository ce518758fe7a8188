#!/bin/bash
# usage: tools/sass.sh <kernel-name-substring> [hist|dump]  -- SASS of one kernel of libdlimgedit.so
LIB=/root/repo/dlimgedit_b200/libdlimgedit.so
cuobjdump -sass $LIB 2>/dev/null | awk -v pat="$1" '
/Function :/ { on = (index($0, pat) > 0); if (on) print; next }
on { print }' > /tmp/sass_$$.txt
if [ "$2" == "dump" ]; then cat /tmp/sass_$$.txt; else
grep -c -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/sass_$$.txt
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/sass_$$.txt | sed -E 's/^\s+\/\*[0-9a-f]{4}\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/;//' | sort | uniq -c | sort -rn | head -${3:-25}
fi
rm -f /tmp/sass_$$.txt
