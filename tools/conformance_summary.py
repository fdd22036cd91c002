"""Summarise a gtest XML report of the reference's templateFuncTests run against the B200 plugin (tools/run_conformance.sh):
per-suite totals and the list of failing tests with the first line of each failure.  Falls back to the text log when the binary
died before it could write the report."""
import collections
import re
import sys
import xml.etree.ElementTree as ET


def from_xml(path):
    rows = collections.OrderedDict()
    failures = []
    root = ET.parse(path).getroot()
    for suite in root.iter("testsuite"):
        name = suite.get("name")
        r = rows.setdefault(name, [0, 0, 0, 0])   # run, passed, failed, skipped
        for case in suite.iter("testcase"):
            status = case.get("status", "run")
            result = case.get("result", "completed")
            fails = case.findall("failure")
            if status != "run" or result == "skipped" or case.find("skipped") is not None:
                r[3] += 1
                continue
            r[0] += 1
            if fails:
                r[2] += 1
                msg = (fails[0].get("message") or "").strip().splitlines()
                failures.append((name + "." + case.get("name"), " | ".join(m.strip() for m in msg[:3])))
            else:
                r[1] += 1
    return rows, failures


def from_log(path):
    rows = collections.OrderedDict()
    failures = []
    for line in open(path, errors="replace"):
        m = re.match(r"\[\s+(OK|FAILED|SKIPPED)\s+\] (\S+?)\.(\S+)", line)
        if not m or "(" not in line and m.group(1) != "SKIPPED":
            continue
        r = rows.setdefault(m.group(2), [0, 0, 0, 0])
        if m.group(1) == "SKIPPED":
            r[3] += 1
        else:
            r[0] += 1
            r[1 if m.group(1) == "OK" else 2] += 1
            if m.group(1) == "FAILED":
                failures.append((m.group(2) + "." + m.group(3).rstrip(","), ""))
    return rows, failures


def main():
    xml, log = sys.argv[1], sys.argv[2]
    try:
        rows, failures = from_xml(xml)
        src = xml
    except Exception as exc:   # no report: the binary crashed or timed out
        rows, failures = from_log(log)
        src = "%s (no XML report: %s)" % (log, exc)
    tot = [sum(r[k] for r in rows.values()) for k in range(4)]
    print("source: %s" % src)
    print("%-110s %6s %6s %6s %6s" % ("suite", "run", "pass", "fail", "skip"))
    for name, r in rows.items():
        print("%-110s %6d %6d %6d %6d" % (name[:110], r[0], r[1], r[2], r[3]))
    print("%-110s %6d %6d %6d %6d" % ("TOTAL", tot[0], tot[1], tot[2], tot[3]))
    if failures:
        print("\nfailing tests (%d):" % len(failures))
        for name, msg in failures:
            print("  %s\n      %s" % (name, msg[:400]))


if __name__ == "__main__":
    main()
