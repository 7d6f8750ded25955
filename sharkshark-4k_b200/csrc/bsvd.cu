// (the BSVD streaming engine lives in bsvd_stream.inc, included by engine.cu)
