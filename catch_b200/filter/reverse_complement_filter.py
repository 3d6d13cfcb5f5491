"""ReverseComplementFilter (host plumbing, catch/filter/reverse_complement_filter.py:18-36): every
input probe followed by its reverse complement, with the headers the reference writes.  No device
work; it is here because the coverage analysis's `rc_too` is driven by the same command-line flag
(bin/design.py:379, :431)."""
from catch_b200.filter.base_filter import BaseFilter


class ReverseComplementFilter(BaseFilter):
    def _filter(self, input):
        output = []
        for p in input:
            p.header = "probe_%s | from target sequence" % p.identifier()
            output.append(p)
            p_rc = p.reverse_complement()
            p_rc.header = "probe_%s | reverse complement of probe_%s" % (p_rc.identifier(), p.identifier())
            output.append(p_rc)
        return output
