"""jammy_flows_b200 -- B200-native (sm_100a) implementation of the jammy_flows hot path.

    import jammy_flows_b200 as jammy_flows
    pdf = jammy_flows.pdf("e4+s2+e4", "gggg+n+gggg").double().cuda()
    log_pdf, log_pdf_base, base = pdf(x)
    x, z, log_pdf, log_pdf_base = pdf.sample(samplesize=1000)

Same constructor / forward / sample / entropy surface as `jammy_flows.pdf` (reference jammy_flows/__init__.py:1,
main/default.py:42-151); the layer math runs in hand-written CUDA kernels behind a C-ABI (include/jammy_b200.h).
"""
from .pdf import pdf  # noqa: F401
from .fully_amortized import fully_amortized_pdf  # noqa: F401

__version__ = "0.1.0"
