//  Swift-on-Linux shim: same type and member names as the reference's Common/SyllableDetectorConfig.swift,
//  Common/SyllableDetector.swift, Common/Resampler.swift and SyllableDetectorCLI/TrackDetector.swift, every call
//  forwarded to libsyldet_cuda.so through the module-mapped C header (include/syldet.h).  Pure forwarding, no arithmetic.
//  UNBUILT here (the image has no Swift toolchain); the identical C-ABI calls are exercised from C++ (cli/) and Python.

import CSyldet
import Foundation

public struct SyllableDetectorConfig {
    public enum Scaling: Int32 { case linear = 0, log = 1, db = 2 }
    public enum ParseError: Error {
        case unableToOpenPath(String)
        case missingValue(String)
        case invalidValue(String)
        case mismatchedLength(String)
    }

    let handle: OpaquePointer

    public var samplingRate: Double { return syldet_config_sampling_rate(handle) }
    public var fourierLength: Int { return Int(syldet_config_fourier_length(handle)) }
    public var windowLength: Int { return Int(syldet_config_window_length(handle)) }
    public var windowOverlap: Int { return Int(syldet_config_window_overlap(handle)) }
    public var timeRange: Int { return Int(syldet_config_time_range(handle)) }
    public var spectrogramScaling: Scaling { return Scaling(rawValue: syldet_config_scaling(handle))! }
    public var freqRange: (Double, Double) {
        var lo = 0.0, hi = 0.0
        syldet_config_freq_range(handle, &lo, &hi)
        return (lo, hi)
    }
    public var thresholds: [Double] {
        var out = [Double](repeating: 0.0, count: Int(syldet_config_threshold_count(handle)))
        syldet_config_thresholds(handle, &out, Int32(out.count))
        return out
    }
    public var netInputs: Int { return Int(syldet_config_net_inputs(handle)) }
    public var netOutputs: Int { return Int(syldet_config_net_outputs(handle)) }

    public init(fromTextFile path: String) throws {
        var h: OpaquePointer? = nil
        let st = syldet_config_load_text(path, &h)
        guard st == SYLDET_OK, let hh = h else {
            let key = String(cString: syldet_config_error_key())
            switch st {
            case SYLDET_ERR_OPEN: throw ParseError.unableToOpenPath(key)
            case SYLDET_ERR_MISSING: throw ParseError.missingValue(key)
            case SYLDET_ERR_INVALID: throw ParseError.invalidValue(key)
            case SYLDET_ERR_MISMATCH: throw ParseError.mismatchedLength(key)
            default: fatalError(String(cString: syldet_last_error()))  // NeuralNet.init / NeuralNetLayer.init fatalError upstream
            }
        }
        handle = hh
    }
}

public final class SyllableDetector {
    public let config: SyllableDetectorConfig
    private var handle: OpaquePointer

    public var lastOutputs: [Float] {
        var out = [Float](repeating: 0.0, count: config.netOutputs)
        syldet_detector_last_outputs(handle, &out, Int32(out.count))
        return out
    }
    public var lastDetected: Bool { return syldet_detector_last_detected(handle) != 0 }

    public init(config: SyllableDetectorConfig, device: Int32 = 0) {
        self.config = config
        var h: OpaquePointer? = nil
        guard syldet_detector_create(config.handle, device, &h) == SYLDET_OK, let hh = h else {
            fatalError(String(cString: syldet_last_error()))  // SyllableDetector.swift:46-60 are fatalError upstream
        }
        handle = hh
    }
    deinit { syldet_detector_destroy(handle) }

    public func appendAudioData(_ data: UnsafeMutablePointer<Float>, withSamples numSamples: Int) {
        if syldet_detector_append(handle, data, Int64(numSamples)) != SYLDET_OK {
            fatalError(String(cString: syldet_last_error()))  // "Insufficient space on buffer." CSTFT.swift:199
        }
    }
    public func processNewValue() -> Bool {
        let r = syldet_detector_process_new_value(handle)
        if r < 0 { fatalError(String(cString: syldet_last_error())) }
        return r == 1
    }
    public func seenSyllable() -> Bool {
        let r = syldet_detector_seen_syllable(handle)
        if r < 0 { fatalError(String(cString: syldet_last_error())) }
        return r == 1
    }
}

/// Stand-in for AVAssetTrack + AVAssetReaderTrackOutput on Linux: hands out Float32 mono buffers at config.samplingRate.
public protocol PCMTrack {
    func copyNextSampleBuffer() -> [Float]?
}

public final class TrackDetector {
    public let track: PCMTrack
    public let detector: SyllableDetector
    public let channel: Int
    public var debounceFrames = 0
    public var debounceTime: Double {
        get { return Double(debounceFrames) / detector.config.samplingRate }
        set { debounceFrames = Int(newValue * detector.config.samplingRate) }
    }
    private var nextOutput: Int
    private var totalSamples = 0
    private var debounceUntil = -1

    public init(track: PCMTrack, config: SyllableDetectorConfig, channel: Int = 0) {
        detector = SyllableDetector(config: config)
        self.track = track
        self.channel = channel
        nextOutput = Int(syldet_config_first_output_sample(config.handle))
    }

    public func process() {
        guard var buffer = track.copyNextSampleBuffer(), 0 < buffer.count else { return }
        let numSamples = buffer.count
        buffer.withUnsafeMutableBufferPointer { detector.appendAudioData($0.baseAddress!, withSamples: numSamples) }
        let thresholds = detector.config.thresholds
        let hop = Int(syldet_config_hop(detector.config.handle))
        while detector.processNewValue() {
            let curOutput = nextOutput
            nextOutput += hop
            let outs = detector.lastOutputs
            var hasDetection = false
            for (i, d) in outs.enumerated() where Double(d) >= thresholds[i] { hasDetection = true; break }
            if hasDetection && debounceUntil < curOutput {
                let curSample = curOutput - totalSamples
                if curSample >= numSamples { fatalError("Unexpected sample number.") }
                print("\(channel),\(curOutput),\(Double(curOutput) / detector.config.samplingRate)", terminator: "")
                for d in outs { print(",\(d)", terminator: "") }
                print("")
                debounceUntil = curOutput + debounceFrames
            }
        }
        totalSamples += numSamples
    }
}

public protocol Resampler {
    func resampleVector(_ data: UnsafePointer<Float>, ofLength numSamples: Int) -> [Float]
}

public final class ResamplerLinear: Resampler {
    private var handle: OpaquePointer
    public init(fromRate samplingRateIn: Double, toRate samplingRateOut: Double) {
        var h: OpaquePointer? = nil
        guard syldet_resampler_linear_create(samplingRateIn, samplingRateOut, &h) == SYLDET_OK, let hh = h else {
            fatalError(String(cString: syldet_last_error()))
        }
        handle = hh
    }
    deinit { syldet_resampler_destroy(handle) }
    public func resampleVector(_ data: UnsafePointer<Float>, ofLength numSamplesIn: Int) -> [Float] {
        var out = [Float](repeating: 0.0, count: Int(syldet_resampler_max_output(handle, Int64(numSamplesIn))))
        var n: Int64 = 0
        if syldet_resampler_process(handle, data, Int64(numSamplesIn), &out, Int64(out.count), &n) != SYLDET_OK {
            fatalError(String(cString: syldet_last_error()))
        }
        out.removeLast(out.count - Int(n))
        return out
    }
}
