timeout 120 python tools/decode_timeline.py uniform
CNB_DECODE_CTAS=148 timeout 120 python tools/decode_timeline.py uniform
