"""FASTA edge-case files shared by tests/test_fasta_native_cpu.py and tests/golden/make_golden_fasta.py."""

CASES = {
    "ka1": b">seq1\nACGTACT\n>seq2\nTTACGTT\n>seq3\nAGATAGA\n",                       # run_graph_builder.py:24-28
    "ka2": b">p1\nACGT\n>p2\nTTAC\n>p3\nAGA",                                            # unit_tests.py:41, no trailing newline
    "crlf_lower_multiline": b">sp|P12345|NAME_HUMAN desc\r\nacde\r\nfghi\r\n\r\n>tr|Q9|x\r\n  klmn  \r\n",
    "lone_cr": b">a\rACD\rEF\r>b\rGG\r",
    "junk_before_header": b"ACGT\n\n  \nXX\n>a\nAC\n",
    "empty_records": b">a\n>b\n\n\n>c\nAC\n>d\n   \n>e",
    "interior_space_tab": b">a\nAC DE\tFG\n>b\n\tAC\x0b\n",
    "bare_header_stops": b">a\nAC\n>b\nDE\n>\nFG\n>c\nHI\n",                            # '>' alone: the reference's generator raises there
    "bare_header_first": b">   \nAC\n>c\nHI\n",
    "pipes": b">||x\nAC\n>|id|\nDE\n>a|\nFG\n",
    "control_ws": b">a\n\x1cAC\x1f\n\x1d\n>b\n\x0cDE\x0c\n",
    "only_text": b"no header at all\nACGT\n",
    "empty": b"",
    "header_with_gt": b">a>b\nAC>DE\n",
}
